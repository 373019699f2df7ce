"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference on CPU.

    python tests/golden/make_golden.py            # needs /root/reference (or $ADFWI_REF)

The reference tree does not exist on the GPU box, so the fixtures are committed.  Every fixture
holds the exact inputs that were fed to the reference's ``forward_kernel`` together with its
outputs (records, illumination maps) and autograd gradients for explicit record cotangents
(loss = sum(W * record)), one gradient set per record component so that every adjoint path is
exercised on its own scale.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_loader  # noqa: E402

torch.set_num_threads(8)


def ricker_integral(nt, dt, f0):
    """Time-integrated Ricker wavelet (what the reference examples feed the propagators)."""
    t = np.arange(nt) * dt
    t0 = 1.2 / f0
    a = (np.pi * f0 * (t - t0)) ** 2
    w = (1 - 2 * a) * np.exp(-a)
    return np.cumsum(w) * dt


def model_2d(nz, nx, rng):
    z = np.linspace(0, 1, nz)[:, None]
    x = np.linspace(0, 1, nx)[None, :]
    vp = 1500 + 1800 * z + 150 * np.sin(2 * np.pi * 2 * x + 3 * z) + 30 * rng.standard_normal((nz, nx))
    return np.clip(vp, 1400, 3600).astype(np.float32)


def acoustic_case(name, free_surface, segments, nz=32, nx=44, nabc=10, nt=150, seed=0):
    ref_loader.load()
    from ADFWI.propagator import acoustic_kernels as ak
    from ADFWI.propagator.boundary_condition import bc_pml
    rng = np.random.default_rng(seed)
    dx = dz = 10.0
    dt = 1e-3
    vp = model_2d(nz, nx, rng)
    rho = (310.0 * vp ** 0.25).astype(np.float32) * (1 + 0.02 * rng.standard_normal((nz, nx))).astype(np.float32)
    damp = bc_pml(nx, nz, dx, dz, pml=nabc, vmax=float(vp.max()), free_surface=False).astype(np.float32)
    src_x = np.array([5, 30], dtype=np.int64)
    src_z = np.array([1, 12], dtype=np.int64)
    ns = len(src_x)
    wav = ricker_integral(nt, dt, 25.0).astype(np.float32)
    src_v = np.stack([wav, 0.7 * np.roll(wav, 7)]).astype(np.float32)
    rcv_x = np.array([0, 9, 17, 17, 26, 43, 22], dtype=np.int64)   # two coincident receivers
    rcv_z = np.array([1, 1, 2, 2, 1, 1, 20], dtype=np.int64)
    nr = len(rcv_x)
    W = {k: rng.standard_normal((ns, nt, nr)).astype(np.float32) for k in ("p", "u", "w")}

    out = {}
    for comp in ("p", "u", "w"):
        v_t = torch.tensor(vp, requires_grad=True)
        r_t = torch.tensor(rho, requires_grad=True)
        rec = ak.forward_kernel(nx, nz, dx, dz, nt, dt, nabc, free_surface,
                                torch.tensor(src_x), torch.tensor(src_z), ns, torch.tensor(src_v),
                                torch.tensor(rcv_x), torch.tensor(rcv_z), nr,
                                torch.tensor(damp), v_t, r_t,
                                checkpoint_segments=segments, device=torch.device("cpu"), dtype=torch.float32)
        loss = (rec[comp] * torch.tensor(W[comp])).sum()
        loss.backward()
        out[f"g_v_{comp}"] = v_t.grad.numpy().copy()
        out[f"g_rho_{comp}"] = r_t.grad.numpy().copy()
    for k in ("p", "u", "w", "forward_wavefield_p", "forward_wavefield_u", "forward_wavefield_w"):
        out[f"rec_{k}"] = rec[k].detach().numpy().copy()
    np.savez_compressed(
        os.path.join(HERE, f"{name}.npz"),
        nz=nz, nx=nx, nabc=nabc, nt=nt, dx=dx, dz=dz, dt=dt, free_surface=free_surface, segments=segments,
        vp=vp, rho=rho, damp=damp, src_x=src_x, src_z=src_z, src_v=src_v, rcv_x=rcv_x, rcv_z=rcv_z,
        W_p=W["p"], W_u=W["u"], W_w=W["w"], **out)
    print(name, {k: float(np.abs(v).max()) for k, v in out.items() if k.startswith("g_")})


def boundary_case():
    ref_loader.load()
    from ADFWI.propagator import boundary_condition as bc
    out = {}
    for fs in (True, False):
        tag = "fs" if fs else "nofs"
        out[f"pml_{tag}"] = bc.bc_pml(23, 17, 10.0, 10.0, pml=7, vmax=3210.5, free_surface=fs)
        bx, bz = bc.bc_pml_xz(23, 17, 10.0, 10.0, pml=7, vmax=3210.5, free_surface=fs)
        out[f"pmlx_{tag}"], out[f"pmlz_{tag}"] = bx, bz
        out[f"gerjan_{tag}"] = bc.bc_gerjan(23, 17, 10.0, 10.0, pml=7, alpha=0.0053, free_surface=fs)
        out[f"sincos_{tag}"] = bc.bc_sincos(23, 17, 10.0, 10.0, pml=7, free_surface=fs)
    np.savez_compressed(os.path.join(HERE, "boundary_profiles.npz"), **out)
    print("boundary_profiles", sorted(out))


if __name__ == "__main__":
    which = sys.argv[1:] or ["acoustic", "boundary", "elastic"]
    if "acoustic" in which:
        acoustic_case("acoustic_fs", True, 2)
        acoustic_case("acoustic_nofs", False, 1, seed=1)
    if "boundary" in which:
        boundary_case()
    if "elastic" in which:
        from make_golden_elastic import main as elastic_main
        elastic_main()
