"""Edge cases of the hot path against the CPU oracle (through the drop-in forward_kernel -> C ABI): colliding receivers (several
receivers in one cell: the adjoint scatter-adds), shots sharing one source cell, one shot / one receiver / a handful of time steps,
sources and receivers on the first and last cells of the physical grid, grids that are not a multiple of the tile size, and a
survey without receivers.  Every fused pipeline (per-step acoustic kernels, whole-sweep acoustic kernels, split-PML and sponge
elastic kernels) and the generic kernels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REC_TOL, GRAD_TOL = 1e-5, 1e-4


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    n = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / n) if n > 0 else float(np.linalg.norm(a - b))


def _acoustic_case(nz, nx, nabc, nt, sx, sz, rx, rz, fs, seed=0, **cfg):
    from adfwi_b200 import synthetic as syn
    from adfwi_b200.propagator import acoustic_kernels as ak
    from adfwi_b200.propagator.boundary_condition import bc_pml
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda:0")
    dx, dt = 10.0, 1e-3
    vp = syn.marmousi_like_vp(nz, nx)
    rho = syn.gardner_rho(vp)
    damp = bc_pml(nx, nz, dx, dx, pml=nabc, vmax=float(vp.max()), free_surface=False).astype(np.float32)
    sx, sz, rx, rz = (np.asarray(a, np.int64) for a in (sx, sz, rx, rz))
    ns, nr = len(sx), len(rx)
    wav = (rng.standard_normal((ns, nt)) * 1e3).astype(np.float32)
    W = [rng.standard_normal((ns, nt, nr)).astype(np.float32) for _ in range(3)]
    t = lambda a: torch.tensor(a, device=dev)
    old = dict(ak.config); ak.config.update(cfg)
    try:
        v = t(vp).requires_grad_(True); r = t(rho).requires_grad_(True)
        rec = ak.forward_kernel(nx, nz, dx, dx, nt, dt, nabc, fs, t(sx), t(sz), ns, t(wav), t(rx), t(rz), nr, t(damp), v, r,
                                checkpoint_segments=1, device=dev)
        if nr > 0:
            sum((rec[k] * t(W[i])).sum() for i, k in enumerate("puw")).backward()
    finally:
        ak.config.clear(); ak.config.update(old)
    coef = O.acoustic_coefficients(vp, rho, damp, dt, dx, nabc, fs)
    ref = O.acoustic_run(coef, nabc, fs, dt, sx, sz, wav, rx, rz, g_rcv=tuple(W) if nr > 0 else None, need_g_alpha2=nr > 0, illum=True)
    for k in "puw":
        got = rec[k].detach().cpu().numpy()
        assert got.shape == (ns, nt, nr)
        if nr > 0:
            assert np.array_equal(got, ref[k]), f"record {k} not bit-identical to the oracle"
    assert rel_l2(rec["forward_wavefield_p"].cpu().numpy(), ref["illum_p"][nabc:nabc + nz, nabc:nabc + nx]) <= 1e-5
    if nr > 0:
        gv, grho = O.acoustic_model_gradients(coef, ref["g_alpha1"], ref["g_alpha2"], dt, dx, nabc)
        assert rel_l2(v.grad.cpu().numpy(), gv) <= GRAD_TOL
        assert rel_l2(r.grad.cpu().numpy(), grho) <= GRAD_TOL


ACOUSTIC_PATHS = [dict(persistent=False), dict(persistent=True), dict(force_generic=True), dict(persistent=False, shots_per_chunk=1)]


@pytest.mark.parametrize("cfg", ACOUSTIC_PATHS)
@pytest.mark.parametrize("fs", [True, False])
def test_acoustic_colliding_receivers_and_shared_source_cells(cfg, fs):
    # three receivers in one cell, two more in the cell next to it; three shots, two of them from the same cell
    _acoustic_case(37, 70, 12, 60, sx=[30, 30, 5], sz=[1, 1, 20], rx=[10, 10, 10, 11, 11, 40], rz=[2, 2, 2, 2, 2, 30], fs=fs, **cfg)


@pytest.mark.parametrize("cfg", ACOUSTIC_PATHS)
def test_acoustic_minimal_survey_and_grid_corners(cfg):
    # one shot, one receiver, three time steps
    _acoustic_case(20, 33, 8, 3, sx=[16], sz=[3], rx=[16], rz=[3], fs=True, **cfg)
    # sources / receivers on the first and last cells of the physical grid; odd grid sizes (ragged tiles)
    _acoustic_case(41, 67, 10, 50, sx=[0, 66], sz=[0, 40], rx=[0, 66, 0, 66], rz=[0, 0, 40, 40], fs=False, **cfg)


@pytest.mark.parametrize("cfg", [dict(persistent=False), dict(persistent=True), dict(force_generic=True)])
def test_acoustic_survey_without_receivers(cfg):
    _acoustic_case(24, 40, 8, 30, sx=[5, 20], sz=[1, 1], rx=[], rz=[], fs=True, **cfg)


def _elastic_case(nz, nx, nabc, nt, sx, sz, rx, rz, fs, abc, order, seed=0, **cfg):
    from adfwi_b200 import synthetic as syn
    from adfwi_b200.propagator import acoustic_kernels as ak, elastic_kernels as ek
    from adfwi_b200.propagator.boundary_condition import bc_gerjan, bc_pml_xz
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda:0")
    dx, dt = 10.0, 1e-3
    vp = syn.marmousi_like_vp(nz, nx)
    model = syn.ElasticGridModel(vp, (vp / np.sqrt(3.0)).astype(np.float32), syn.gardner_rho(vp), dx=dx, dz=dx, nabc=nabc, free_surface=fs,
                                 abc_type=abc, requires_grad=(), device=dev)
    model.forward()
    planes = ("C11", "C13", "C33", "C55", "bx", "bz")
    idx = {"C11": 0, "C13": 2, "C33": 11, "C55": 18}
    L = {k: (model.CC[idx[k]] if k in idx else getattr(model, k)).detach().clone().requires_grad_(True) for k in planes}
    CC = list(model.CC)
    for k, i in idx.items():
        CC[i] = L[k]
    pml = abc == "PML"
    if pml:
        bcx, bcz = bc_pml_xz(nx, nz, dx, dx, pml=nabc, vmax=float(vp.max()), free_surface=fs)
        bc = dict(bcx=np.asarray(bcx, np.float32), bcz=np.asarray(bcz, np.float32))
    else:
        bc = dict(damp=np.asarray(bc_gerjan(nx, nz, dx, dx, pml=nabc, alpha=0.0053, free_surface=fs), np.float32))
    sx, sz, rx, rz = (np.asarray(a, np.int64) for a in (sx, sz, rx, rz))
    ns, nr = len(sx), len(rx)
    wav = (rng.standard_normal((ns, nt)) * 1e3).astype(np.float32)
    mt = np.broadcast_to(np.eye(3, dtype=np.float32), (ns, 3, 3)).copy()
    mt[:, 0, 2] = 0.4; mt[:, 2, 0] = 0.4
    comps = ("txx", "tzz", "txz", "vx", "vz")
    W = [rng.standard_normal((ns, nt, nr)).astype(np.float32) for _ in comps]
    t = lambda a: torch.tensor(a, device=dev)
    old = dict(ak.config); ak.config.update(cfg)
    try:
        rec = ek.forward_kernel(nx, nz, dx, dx, nt, dt, nabc, fs, t(sx), t(sz), ns, t(wav), t(mt), t(rx), t(rz), nr, abc,
                                t(bc["bcx"]) if pml else None, t(bc["bcz"]) if pml else None, None if pml else t(bc["damp"]), None, None,
                                L["bx"], L["bz"], CC, fd_order=order, n_segments=1, device=dev)
        sum((rec[k] * t(W[i])).sum() for i, k in enumerate(comps)).backward()
    finally:
        ak.config.clear(); ak.config.update(old)
    ref = O.elastic_run({k: L[k].detach().cpu().numpy() for k in planes}, abc, order, fs, nz, nx, nabc, dx, dx, dt, sx, sz, wav, mt, rx, rz,
                        g_rcv=W, **bc)
    for k in comps:
        assert np.array_equal(rec[k].detach().cpu().numpy(), ref[k]), f"record {k} not bit-identical to the oracle"
    for k in planes:
        assert rel_l2(L[k].grad.cpu().numpy(), ref["g_own"][k]) <= GRAD_TOL, k


@pytest.mark.parametrize("abc", ["PML", "gerjan"])
@pytest.mark.parametrize("order", [4, 6])
@pytest.mark.parametrize("cfg", [dict(), dict(force_generic=True), dict(shots_per_chunk=1)])
def test_elastic_colliding_receivers_shared_sources_and_corners(abc, order, cfg):
    # three receivers in one cell, two shots from one cell, a source and receivers on the last physical cells, ragged tiles
    _elastic_case(29, 75, 10, 40, sx=[30, 30, 74], sz=[2, 2, 28], rx=[12, 12, 12, 13, 0, 74], rz=[3, 3, 3, 3, 0, 28], fs=True, abc=abc,
                  order=order, **cfg)


@pytest.mark.parametrize("abc", ["PML", "gerjan"])
def test_elastic_minimal_survey(abc):
    _elastic_case(18, 34, 6, 3, sx=[17], sz=[5], rx=[17], rz=[5], fs=False, abc=abc, order=4)
