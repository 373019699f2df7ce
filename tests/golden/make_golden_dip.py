"""Golden fixture for the autograd contract of SURVEY.md 8(b): `v` handed to forward_kernel is a NON-LEAF tensor (the output
of a small convolutional generator, as in the reference's deep-image-prior models, ADFWI/dip/dip_acoustic_model.py:161),
and the gradients of a waveform misfit must reach the generator's weights.  Runs the UNMODIFIED reference kernel on CPU.

    python tests/golden/make_golden_dip.py        # needs /root/reference (or $ADFWI_REF)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_loader  # noqa: E402

torch.set_num_threads(8)


def make_net():
    """4 -> 8 -> 1 channel 3x3 generator; vp = 2500 + 800 tanh(net(z))."""
    return torch.nn.Sequential(torch.nn.Conv2d(4, 8, 3, padding=1), torch.nn.Tanh(), torch.nn.Conv2d(8, 1, 3, padding=1))


def main():
    ref_loader.load()
    from ADFWI.propagator import acoustic_kernels as ak
    from ADFWI.propagator.boundary_condition import bc_pml
    torch.manual_seed(11)
    rng = np.random.default_rng(11)
    nz, nx, nabc, nt = 30, 46, 10, 180
    dx = dz = 10.0
    dt = 1e-3
    net = make_net()
    z = torch.tensor(rng.standard_normal((1, 4, nz, nx)).astype(np.float32))
    t = np.arange(nt) * dt
    a = (np.pi * 25.0 * (t - 1.2 / 25.0)) ** 2
    wav = (np.cumsum((1 - 2 * a) * np.exp(-a)) * dt).astype(np.float32)
    src_x = torch.tensor([7, 23, 38]); src_z = torch.tensor([1, 1, 1])
    rcv_x = torch.arange(0, nx, 3); rcv_z = torch.full_like(rcv_x, 1)
    src_v = torch.tensor(np.stack([wav] * 3))
    W = torch.tensor(rng.standard_normal((3, nt, len(rcv_x))).astype(np.float32))
    vp = 2500.0 + 800.0 * torch.tanh(net(z))[0, 0]
    rho = 310.0 * vp.detach() ** 0.25
    damp = torch.tensor(bc_pml(nx, nz, dx, dz, pml=nabc, vmax=3300.0, free_surface=False).astype(np.float32))
    rec = ak.forward_kernel(nx, nz, dx, dz, nt, dt, nabc, True, src_x, src_z, 3, src_v, rcv_x, rcv_z, len(rcv_x), damp, vp, rho,
                            checkpoint_segments=1, device="cpu", dtype=torch.float32)
    loss = (rec["p"] * W).sum()
    loss.backward()
    out = dict(nz=nz, nx=nx, nabc=nabc, nt=nt, dx=dx, dz=dz, dt=dt, z=z.numpy(), src_x=src_x.numpy(), src_z=src_z.numpy(), src_v=src_v.numpy(),
               rcv_x=rcv_x.numpy(), rcv_z=rcv_z.numpy(), W=W.numpy(), damp=damp.numpy(), rho=rho.numpy(), vp=vp.detach().numpy(),
               rec_p=rec["p"].detach().numpy(), loss=float(loss))
    for k, p in net.state_dict().items():
        out["w_" + k] = p.numpy()
    for k, p in net.named_parameters():
        out["g_" + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "dip_acoustic.npz"), **out)
    print("loss", float(loss), {k: float(np.abs(v).max()) for k, v in out.items() if k.startswith("g_")})


if __name__ == "__main__":
    main()
