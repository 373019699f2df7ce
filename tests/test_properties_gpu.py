"""Size-independent properties of the CUDA path on the FULL BASELINE.json grids (C2: iso-acoustic
350x1700, C3: iso-elastic 350x1700 split-PML O(2,4) with free surface), where the CPU oracle would take
minutes: a slice of a few shots x a few hundred steps is checked through

  * the adjoint identity  <F s, w> == <s, F^T w>  (the time loop is linear in the source wavelet; F^T w
    is what the hand-written reverse kernels return as the source gradient),
  * linearity of the records in the source,
  * shot independence (a shot's records do not depend on its batch: bit-exact; the gradient of a batch
    is the sum of the per-shot gradients),
  * invariance to the time-axis memory scheme (store-all vs checkpoint + recompute): records bit-exact,
    gradients to round-off,
  * a zero source gives exactly zero records and a zero gradient.

All calls go through the drop-in forward_kernel -> C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NZ, NX, NABC = 350, 1700, 50
DX = 10.0
DT = 1e-3


def rel(a, b):
    a = a.double(); b = b.double()
    return float((a - b).norm() / b.norm())


def _cfg(**kw):
    from adfwi_b200.propagator import acoustic_kernels as ak
    old = dict(ak.config)
    ak.config.update(kw)
    return ak, old


def _acoustic_setup(ns, nt, nr=1700, seed=0):
    from adfwi_b200 import synthetic as syn
    from adfwi_b200.propagator.boundary_condition import bc_pml
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(seed)
    vp = syn.marmousi_like_vp(NZ, NX)
    rho = syn.gardner_rho(vp)
    damp = torch.tensor(bc_pml(NX, NZ, DX, DX, NABC, float(vp.max()), free_surface=False), dtype=torch.float32, device=dev)
    sx = torch.linspace(2, NX - 3, ns).round().long().to(dev); sz = torch.full((ns,), 1, device=dev).long()
    rx = torch.linspace(0, NX - 1, nr).round().long().to(dev); rz = torch.full((nr,), 1, device=dev).long()
    src = torch.randn(ns, nt, generator=g).to(dev)
    return dict(dev=dev, vp=torch.tensor(vp, device=dev), rho=torch.tensor(rho, device=dev), damp=damp,
                sx=sx, sz=sz, rx=rx, rz=rz, src=src, ns=ns, nt=nt, nr=nr, gen=g)


def _acoustic_forward(S, src, vp=None, sel=None):
    from adfwi_b200.propagator import acoustic_kernels as ak
    sx, sz = (S["sx"], S["sz"]) if sel is None else (S["sx"][sel], S["sz"][sel])
    vp = S["vp"] if vp is None else vp
    return ak.forward_kernel(NX, NZ, DX, DX, src.shape[1], DT, NABC, True, sx, sz, src.shape[0], src, S["rx"], S["rz"], S["nr"],
                             S["damp"], vp, S["rho"], checkpoint_segments=1, device=S["dev"])


def test_acoustic_full_grid_adjoint_identity_and_linearity():
    S = _acoustic_setup(ns=3, nt=300)
    g = S["gen"]
    s1 = S["src"].clone().requires_grad_(True)
    s2 = torch.randn(S["ns"], S["nt"], generator=g).to(S["dev"])
    W = {k: torch.randn(S["ns"], S["nt"], S["nr"], generator=g).to(S["dev"]) for k in ("p", "u", "w")}
    # u, w records are ~1e-6 of p (dt/(rho dz)): scale their weights up so that all three matter in the identity
    scale = {"p": 1.0, "u": 1e6, "w": 1e6}
    r1 = _acoustic_forward(S, s1)
    sum((r1[k] * W[k]).sum() * scale[k] for k in W).backward()
    FTw = s1.grad.double()
    with torch.no_grad():
        r2 = _acoustic_forward(S, s2)
        lhs = sum((r2[k].double() * W[k].double()).sum() * scale[k] for k in W)
        rhs = (s2.double() * FTw).sum()
        assert abs(float(lhs - rhs)) <= 1e-4 * max(abs(float(lhs)), abs(float(rhs))), (float(lhs), float(rhs))
        a, b = 0.75, -1.5
        r3 = _acoustic_forward(S, a * s1.detach() + b * s2)
        for k in W:
            assert rel(r3[k], a * r1[k].detach() + b * r2[k]) <= 1e-5, k


def test_acoustic_full_grid_shot_independence_and_memory_scheme():
    S = _acoustic_setup(ns=4, nt=200)
    g = S["gen"]
    W = torch.randn(S["ns"], S["nt"], S["nr"], generator=g).to(S["dev"])
    out = {}
    for tag, cfg in (("store_all", dict(ckpt_interval=None)), ("ckpt", dict(ckpt_interval=48, shots_per_group=3))):
        ak, old = _cfg(**cfg)
        try:
            vp = S["vp"].clone().requires_grad_(True)
            r = _acoustic_forward(S, S["src"], vp=vp)
            (r["p"] * W).sum().backward()
            out[tag] = (r["p"].detach(), vp.grad.clone())
        finally:
            ak.config.clear(); ak.config.update(old)
    assert torch.equal(out["store_all"][0], out["ckpt"][0])
    assert rel(out["ckpt"][1], out["store_all"][1]) <= 1e-5
    gsum = torch.zeros_like(S["vp"])
    for i in range(S["ns"]):
        vp = S["vp"].clone().requires_grad_(True)
        sel = torch.tensor([i], device=S["dev"])
        r = _acoustic_forward(S, S["src"][i:i + 1], vp=vp, sel=sel)
        assert torch.equal(r["p"].detach()[0], out["store_all"][0][i]), i
        (r["p"] * W[i:i + 1]).sum().backward()
        gsum += vp.grad
    assert rel(gsum, out["store_all"][1]) <= 1e-5


def test_acoustic_zero_source_is_exactly_zero():
    S = _acoustic_setup(ns=2, nt=60)
    vp = S["vp"].clone().requires_grad_(True)
    r = _acoustic_forward(S, torch.zeros_like(S["src"]), vp=vp)
    for k in ("p", "u", "w"):
        assert not bool(r[k].detach().any())
    r["p"].sum().backward()
    assert not bool(vp.grad.any())


# ---- elastic, C3 grid -------------------------------------------------------------------------------
def _elastic_setup(ns, nt, nr=1700, seed=0, fd_order=4):
    from adfwi_b200 import synthetic as syn
    from adfwi_b200.propagator.boundary_condition import bc_pml_xz
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(seed)
    vp = syn.marmousi_like_vp(NZ, NX)
    model = syn.ElasticGridModel(vp, vp / np.sqrt(3.0), syn.gardner_rho(vp), dx=DX, dz=DX, nabc=NABC, free_surface=True,
                                 abc_type="PML", requires_grad=(), device=dev)
    model.forward()
    bcx, bcz = bc_pml_xz(NX, NZ, DX, DX, pml=NABC, vmax=float(vp.max()), free_surface=True)
    sx = torch.linspace(2, NX - 3, ns).round().long().to(dev); sz = torch.full((ns,), 10, device=dev).long()
    rx = torch.linspace(0, NX - 1, nr).round().long().to(dev); rz = torch.full((nr,), 10, device=dev).long()
    mt = torch.eye(3).repeat(ns, 1, 1).to(dev)
    mt[:, 0, 2] = 0.3; mt[:, 2, 0] = 0.3
    return dict(dev=dev, model=model, bcx=torch.tensor(bcx, dtype=torch.float32, device=dev), bcz=torch.tensor(bcz, dtype=torch.float32, device=dev),
                sx=sx, sz=sz, rx=rx, rz=rz, mt=mt, src=torch.randn(ns, nt, generator=g).to(dev), ns=ns, nt=nt, nr=nr, gen=g, order=fd_order)


COEF = {"C11": 0, "C13": 2, "C33": 11, "C55": 18}


def _elastic_forward(S, src, leaves=None, sel=None):
    from adfwi_b200.propagator import elastic_kernels as ek
    m = S["model"]
    CC = list(m.CC)
    bx, bz = m.bx, m.bz
    if leaves is not None:
        for k, i in COEF.items():
            CC[i] = leaves[k]
        bx, bz = leaves["bx"], leaves["bz"]
    sx, sz, mt = (S["sx"], S["sz"], S["mt"]) if sel is None else (S["sx"][sel], S["sz"][sel], S["mt"][sel])
    return ek.forward_kernel(NX, NZ, DX, DX, src.shape[1], DT, NABC, True, sx, sz, src.shape[0], src, mt, S["rx"], S["rz"], S["nr"],
                             "PML", S["bcx"], S["bcz"], None, None, None, bx, bz, CC, fd_order=S["order"], n_segments=1, device=S["dev"])


def _leaves(S):
    m = S["model"]
    L = {k: m.CC[i].detach().clone().requires_grad_(True) for k, i in COEF.items()}
    L["bx"] = m.bx.detach().clone().requires_grad_(True); L["bz"] = m.bz.detach().clone().requires_grad_(True)
    return L


REC = ("txx", "tzz", "txz", "vx", "vz")
# stresses are ~rho*vp (1e7) times the velocities: weight them down so that all five records matter
RSCALE = {"txx": 1e-7, "tzz": 1e-7, "txz": 1e-7, "vx": 1.0, "vz": 1.0}


@pytest.mark.parametrize("order", [4, 6])
def test_elastic_full_grid_adjoint_identity_and_linearity(order):
    S = _elastic_setup(ns=2, nt=200, fd_order=order)
    g = S["gen"]
    s1 = S["src"].clone().requires_grad_(True)
    s2 = torch.randn(S["ns"], S["nt"], generator=g).to(S["dev"])
    W = {k: torch.randn(S["ns"], S["nt"], S["nr"], generator=g).to(S["dev"]) for k in REC}
    r1 = _elastic_forward(S, s1)
    sum((r1[k] * W[k]).sum() * RSCALE[k] for k in REC).backward()
    FTw = s1.grad.double()
    with torch.no_grad():
        r2 = _elastic_forward(S, s2)
        lhs = sum((r2[k].double() * W[k].double()).sum() * RSCALE[k] for k in REC)
        rhs = (s2.double() * FTw).sum()
        assert abs(float(lhs - rhs)) <= 1e-4 * max(abs(float(lhs)), abs(float(rhs))), (float(lhs), float(rhs))
        a, b = 0.75, -1.5
        r3 = _elastic_forward(S, a * s1.detach() + b * s2)
        for k in REC:
            assert rel(r3[k], a * r1[k].detach() + b * r2[k]) <= 1e-5, k


def test_elastic_full_grid_shot_independence_and_memory_scheme():
    S = _elastic_setup(ns=3, nt=120)
    g = S["gen"]
    W = {k: torch.randn(S["ns"], S["nt"], S["nr"], generator=g).to(S["dev"]) for k in ("vx", "vz")}
    out = {}
    for tag, cfg in (("store_all", dict(ckpt_interval=None)), ("ckpt", dict(ckpt_interval=32, shots_per_group=2))):
        ak, old = _cfg(**cfg)
        try:
            L = _leaves(S)
            r = _elastic_forward(S, S["src"], leaves=L)
            sum((r[k] * W[k]).sum() for k in W).backward()
            out[tag] = ({k: r[k].detach() for k in REC}, {k: v.grad.clone() for k, v in L.items()})
        finally:
            ak.config.clear(); ak.config.update(old)
    for k in REC:
        assert torch.equal(out["store_all"][0][k], out["ckpt"][0][k]), k
    for k in out["ckpt"][1]:
        assert rel(out["ckpt"][1][k], out["store_all"][1][k]) <= 1e-5, k
    gsum = None
    for i in range(S["ns"]):
        L = _leaves(S)
        sel = torch.tensor([i], device=S["dev"])
        r = _elastic_forward(S, S["src"][i:i + 1], leaves=L, sel=sel)
        for k in REC:
            assert torch.equal(r[k].detach()[0], out["store_all"][0][k][i]), (i, k)
        sum((r[k] * W[k][i:i + 1]).sum() for k in W).backward()
        gsum = {k: v.grad.clone() for k, v in L.items()} if gsum is None else {k: gsum[k] + v.grad for k, v in L.items()}
    for k in gsum:
        assert rel(gsum[k], out["store_all"][1][k]) <= 1e-5, k


def test_elastic_zero_source_is_exactly_zero():
    S = _elastic_setup(ns=2, nt=40)
    L = _leaves(S)
    r = _elastic_forward(S, torch.zeros_like(S["src"]), leaves=L)
    for k in REC:
        assert not bool(r[k].detach().any())
    (r["vx"].sum() + r["vz"].sum()).backward()
    for k, v in L.items():
        assert not bool(v.grad.any()), k
