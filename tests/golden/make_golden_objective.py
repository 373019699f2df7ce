"""Fixtures for the record post-processing and the regularisers from the UNMODIFIED reference classes (CPU autograd):

    python tests/golden/make_golden_objective.py        # needs /root/reference (or $ADFWI_REF)

objective_misfit.npz:  records syn / obs (3 shots x 240 samples x 9 traces, band-limited, with a dead-quiet early part and
                       one trace pair that is nearly identical), and for kind in {L2, global correlation} x normalize in {0, 1}:
                       loss and d loss / d syn as `Misfit.forward` + the driver's normalisation + loss.backward() produce them.
objective_regularization.npz: a model plane and value / gradient of TV_1order, Tikhonov_1order, TV_2order, Tikhonov_2order."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_loader  # noqa: E402


def records(rng, ns, nt, nr):
    t = np.arange(nt)[None, :, None]
    f = 0.02 + 0.03 * rng.random((ns, 1, nr)); ph = 6.28 * rng.random((ns, 1, nr)); t0 = 40 + 120 * rng.random((ns, 1, nr))
    w = np.exp(-((t - t0) / 25.0) ** 2) * np.sin(6.28 * f * (t - t0) + ph) * (0.5 + rng.random((ns, 1, nr)))
    return (w + 1e-3 * rng.standard_normal((ns, nt, nr))).astype(np.float32)


def main():
    ref_loader.load()
    from ADFWI.fwi.misfit import Misfit_waveform_L2, Misfit_global_correlation
    from ADFWI.fwi.regularization import regularization_TV_1order as TV_1order, regularization_Tikhonov_1order as Tikhonov_1order, regularization_TV_2order as TV_2order, regularization_Tikhonov_2order as Tikhonov_2order
    rng = np.random.default_rng(7)
    ns, nt, nr = 3, 240, 9
    syn = records(rng, ns, nt, nr) * 3.7e-3            # physical amplitudes are far from 1
    obs = records(rng, ns, nt, nr)
    obs[0, :, 2] = (syn[0, :, 2] / np.abs(syn[0, :, 2]).max() * 1.001 + 1e-4 * rng.standard_normal(nt)).astype(np.float32)   # small residual
    obs = (obs / np.abs(obs).max(axis=1, keepdims=True)).astype(np.float32)     # the drivers normalise obs once (acoustic_fwi.py:68-70)
    out = dict(syn=syn, obs=obs)
    for kname, cls in (("l2", Misfit_waveform_L2), ("gc", Misfit_global_correlation)):
        for norm in (0, 1):
            for dt in (1.0, 2e-3):
                s = torch.tensor(syn, requires_grad=True)
                y = s / (torch.max(torch.abs(s), axis=1, keepdim=True).values) if norm else s
                loss = cls(dt=dt).forward(torch.tensor(obs), y)
                loss.backward()
                tag = f"{kname}_n{norm}_dt{'1' if dt == 1.0 else 's'}"
                out["loss_" + tag] = np.float64(loss.item()); out["g_" + tag] = s.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "objective_misfit.npz"), dt_s=2e-3, **out)
    print({k: float(v) for k, v in out.items() if k.startswith("loss")})

    nz, nx = 37, 52
    z = np.linspace(0, 1, nz)[:, None]; x = np.linspace(0, 1, nx)[None, :]
    m = (1500 + 1800 * z + 200 * np.sin(7 * x + 3 * z) + 20 * rng.standard_normal((nz, nx))).astype(np.float32)
    m[10:14, 20:30] = 2500.0                              # flat patch: zero differences (sign(0) = 0 in the TV gradient)
    reg = dict(m=m, dx=10.0, dz=12.5, alphax=3e-4, alphaz=7e-4)
    for kind, cls in enumerate((TV_1order, Tikhonov_1order, TV_2order, Tikhonov_2order)):
        mt = torch.tensor(m, requires_grad=True)
        r = cls(nx, nz, 10.0, 12.5, 3e-4, 7e-4, step_size=1000, gamma=1)
        v = r.forward(mt)
        v.backward()
        reg[f"value_{kind}"] = np.float64(v.item()); reg[f"g_{kind}"] = mt.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "objective_regularization.npz"), **reg)
    print({k: float(v) for k, v in reg.items() if k.startswith("value")})


if __name__ == "__main__":
    main()
