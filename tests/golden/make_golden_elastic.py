"""Elastic golden fixtures: the UNMODIFIED reference's elastic forward_kernel on CPU (see
make_golden.py).  The coefficient planes are produced by the reference's own parameterisation
chain (ADFWI/model/parameters.py) from leaf vp, vs, rho, eps, delta, so the fixtures pin both the
kernel-level gradients (w.r.t. C11,C13,C33,C55,bx,bz incl. their ragged shapes) and the
model-level gradients (w.r.t. vp, vs, rho, eps, delta)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_loader  # noqa: E402


def elastic_case(name, abc, order, free_surface, segments, nz=30, nx=40, nabc=8, nt=80, seed=0):
    ref_loader.load()
    from ADFWI.propagator import elastic_kernels as ek
    from ADFWI.propagator import boundary_condition as bc
    from ADFWI.model import parameters as P
    from make_golden import model_2d, ricker_integral
    rng = np.random.default_rng(seed)
    dx = dz = 10.0
    dt = 1e-3
    vp = model_2d(nz, nx, rng)
    vs = (vp / np.sqrt(3.0) * (1 + 0.03 * rng.standard_normal((nz, nx)))).astype(np.float32)
    rho = (310.0 * vp ** 0.25).astype(np.float32)
    eps = (0.1 + 0.05 * rng.random((nz, nx))).astype(np.float32)
    delta = (-0.05 + 0.1 * rng.random((nz, nx))).astype(np.float32)
    gamma = np.zeros((nz, nx), np.float32)
    if abc == "PML":
        bcx, bcz = bc.bc_pml_xz(nx, nz, dx, dz, pml=nabc, vmax=float(vp.max()), free_surface=free_surface)
        bcx_t, bcz_t, damp_t = torch.tensor(bcx, dtype=torch.float32), torch.tensor(bcz, dtype=torch.float32), None
    else:
        damp = bc.bc_gerjan(nx, nz, dx, dz, pml=nabc, alpha=0.0053, free_surface=free_surface)
        bcx_t, bcz_t, damp_t = None, None, torch.tensor(damp, dtype=torch.float32)
    src_x = np.array([6, 29], dtype=np.int64); src_z = np.array([4, 14], dtype=np.int64)
    ns = 2
    wav = ricker_integral(nt, dt, 30.0).astype(np.float32)
    src_v = np.stack([wav, 0.6 * np.roll(wav, 5)]).astype(np.float32)
    mt = np.zeros((ns, 3, 3), np.float32)
    mt[0] = np.eye(3); mt[1] = np.array([[0.8, 0, 0.3], [0, 0, 0], [0.3, 0, -0.5]])
    rcv_x = np.array([1, 11, 20, 20, 38], dtype=np.int64); rcv_z = np.array([2, 2, 5, 5, 17], dtype=np.int64)
    nr = len(rcv_x)
    comps = ("txx", "tzz", "txz", "vx", "vz")
    W = {k: rng.standard_normal((ns, nt, nr)).astype(np.float32) for k in comps}
    out = {}
    for tag, use in (("stress", ("txx", "tzz", "txz")), ("vel", ("vx", "vz"))):
        leaves = {k: torch.tensor(v, requires_grad=True) for k, v in dict(vp=vp, vs=vs, rho=rho, eps=eps, delta=delta).items()}
        gam = torch.tensor(gamma)
        mu, lamu, lam, b = P.vs_vp_to_Lame(leaves["vp"], leaves["vs"], leaves["rho"])
        C11, C13, C33, C44, C66 = P.thomsen_to_elastic_moduli(leaves["vp"], leaves["vs"], leaves["rho"], leaves["eps"], leaves["delta"], gam)
        zero = torch.zeros((nz, nx))
        CC = [C11, zero, C13, zero, zero, zero, zero, zero, zero, zero, zero, C33, zero, zero, zero, C44, zero, zero, zero, zero, C66]
        CC = P.elastic_moduli_for_TI(CC, anisotropic_type="VTI")
        C55 = CC[18]
        bx, bz, muxz, C44s, C55s, C66s = P.parameter_staggered_grid(mu, b, C44, C55, C66, nx, nz)
        CC[15], CC[18], CC[20] = C44s, C55s, C66s
        # the kernel inputs become separate graph nodes (clones) so that their retained gradients
        # are the kernel-level ones (C33 also feeds C11 and C13 upstream)
        for idx in (0, 2, 11, 18):
            CC[idx] = CC[idx].clone()
        bx, bz = bx.clone(), bz.clone()
        planes = dict(C11=CC[0], C13=CC[2], C33=CC[11], C55=CC[18], bx=bx, bz=bz)
        for t in planes.values():
            t.retain_grad()
        rec = ek.forward_kernel(nx, nz, dx, dz, nt, dt, nabc, free_surface,
                                torch.tensor(src_x), torch.tensor(src_z), ns, torch.tensor(src_v), torch.tensor(mt),
                                torch.tensor(rcv_x), torch.tensor(rcv_z), nr,
                                abc, bcx_t, bcz_t, damp_t, lamu, lam, bx, bz, CC,
                                fd_order=order, n_segments=segments, device=torch.device("cpu"), dtype=torch.float32)
        loss = sum((rec[k] * torch.tensor(W[k])).sum() for k in use)
        loss.backward()
        for k, t in leaves.items():
            out[f"g_{k}_{tag}"] = t.grad.numpy().copy()
        for k, t in planes.items():
            out[f"g_{k}_{tag}"] = t.grad.numpy().copy()
    for k in comps:
        out[f"rec_{k}"] = rec[k].detach().numpy().copy()
        out[f"fw_{k}"] = rec[f"forward_wavefield_{k}"].detach().numpy().copy()
    for k, t in planes.items():
        out[f"in_{k}"] = t.detach().numpy().copy()
    extra = dict(bcx=bcx_t.numpy(), bcz=bcz_t.numpy()) if abc == "PML" else dict(damp=damp_t.numpy())
    np.savez_compressed(
        os.path.join(HERE, f"{name}.npz"),
        nz=nz, nx=nx, nabc=nabc, nt=nt, dx=dx, dz=dz, dt=dt, free_surface=free_surface, segments=segments,
        abc=abc, order=order, vp=vp, vs=vs, rho=rho, eps=eps, delta=delta,
        src_x=src_x, src_z=src_z, src_v=src_v, mt=mt, rcv_x=rcv_x, rcv_z=rcv_z,
        **{f"W_{k}": W[k] for k in comps}, **extra, **out)
    print(name, {k: float(np.abs(v).max()) for k, v in out.items() if k.startswith("g_v") or k.startswith("g_C11")})


CASES = [(abc, order, fs) for abc in ("PML", "gerjan") for order in (4, 6) for fs in (True, False)]


def case_name(abc, order, fs):
    return f"elastic_{abc.lower()}_o{order}_{'fs' if fs else 'nofs'}"


def main():
    for i, (abc, order, fs) in enumerate(CASES):
        elastic_case(case_name(abc, order, fs), abc, order, fs, segments=1 + (i % 3), seed=10 + i)


if __name__ == "__main__":
    main()
