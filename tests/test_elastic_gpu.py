"""GPU parity of the elastic path (through the C ABI via the drop-in forward_kernel) against the
committed golden fixtures of the unmodified reference: all eight variants
{split-PML, Cerjan sponge} x {O(2,4), O(2,6)} x {free surface on, off}."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REC_TOL = 1e-5     # BASELINE.json: synthetic shot records rel-L2 <= 1e-5
GRAD_TOL = 1e-4    # BASELINE.json: gradients rel-L2 <= 1e-4
COMPS = ("txx", "tzz", "txz", "vx", "vz")
PLANES = ("C11", "C13", "C33", "C55", "bx", "bz")
CASES = [f"elastic_{abc}_o{o}_{fs}" for abc in ("pml", "gerjan") for o in (4, 6) for fs in ("fs", "nofs")]


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _run(g, use, **cfg):
    from adfwi_b200.propagator import acoustic_kernels as ak, elastic_kernels as ek
    old = dict(ak.config)
    ak.config.update(cfg)
    try:
        dev = torch.device("cuda:0")
        t = lambda k: torch.tensor(g[k], device=dev)
        nz, nx = int(g["nz"]), int(g["nx"])
        leaves = {k: t("in_" + k).requires_grad_(True) for k in PLANES}
        zero = torch.zeros((nz, nx), device=dev)
        CC = [zero] * 21
        CC[0], CC[2], CC[11], CC[18] = leaves["C11"], leaves["C13"], leaves["C33"], leaves["C55"]
        abc = str(g["abc"])
        bcx = t("bcx") if abc == "PML" else None
        bcz = t("bcz") if abc == "PML" else None
        damp = None if abc == "PML" else t("damp")
        rec = ek.forward_kernel(nx, nz, float(g["dx"]), float(g["dz"]), int(g["nt"]), float(g["dt"]), int(g["nabc"]),
                                bool(g["free_surface"]), t("src_x"), t("src_z"), len(g["src_x"]), t("src_v"), t("mt"),
                                t("rcv_x"), t("rcv_z"), len(g["rcv_x"]), abc, bcx, bcz, damp, None, None,
                                leaves["bx"], leaves["bz"], CC, fd_order=int(g["order"]), n_segments=int(g["segments"]),
                                device=dev, dtype=torch.float32)
        sum((rec[k] * t("W_" + k)).sum() for k in use).backward()
        return rec, {k: v.grad.cpu().numpy() for k, v in leaves.items()}
    finally:
        ak.config.clear(); ak.config.update(old)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("cfg", [dict(), dict(ckpt_interval=30, shots_per_group=1), dict(force_generic=True),
                                 dict(ckpt_interval=25, shots_per_chunk=1)])
def test_golden_records_and_gradients(golden_dir, name, cfg):
    g = np.load(f"{golden_dir}/{name}.npz")
    for tag, use in (("stress", ("txx", "tzz", "txz")), ("vel", ("vx", "vz"))):
        rec, grads = _run(g, use, **cfg)
        for k in COMPS:
            got = rec[k].detach().cpu().numpy()
            assert rel_l2(got, g["rec_" + k]) <= REC_TOL, (name, k)
            assert np.array_equal(got, g["rec_" + k]), f"{name}: record {k} not bit-identical to the reference"
        for k in PLANES:
            e = rel_l2(grads[k], g[f"g_{k}_{tag}"])
            assert e <= GRAD_TOL and e <= 2e-5, (name, tag, k, e)
    for k in COMPS:
        assert rel_l2(rec["forward_wavefield_" + k].cpu().numpy(), g["fw_" + k]) <= 1e-5, (name, k)


@pytest.mark.parametrize("name", ["elastic_pml_o4_fs", "elastic_gerjan_o6_nofs"])
def test_model_level_gradients(golden_dir, name):
    """vp / vs / rho / eps / delta gradients through ElasticPropagator + the torch parameterisation."""
    from adfwi_b200 import synthetic as syn
    from adfwi_b200.propagator import ElasticPropagator
    g = np.load(f"{golden_dir}/{name}.npz")
    dev = torch.device("cuda:0")
    abc = str(g["abc"])
    model = syn.ElasticGridModel(g["vp"], g["vs"], g["rho"], eps=g["eps"], delta=g["delta"], dx=float(g["dx"]), dz=float(g["dz"]),
                                 nabc=int(g["nabc"]), free_surface=bool(g["free_surface"]), abc_type=abc,
                                 requires_grad=("vp", "vs", "rho", "eps", "delta"), device=dev)
    src = syn.Source(np.stack([g["src_x"], g["src_z"]], 1), g["src_v"], int(g["nt"]), float(g["dt"]), 30.0, moment_tensor=g["mt"])
    rcv = syn.Receiver(np.stack([g["rcv_x"], g["rcv_z"]], 1))
    prop = ElasticPropagator(model, syn.Survey(src, rcv), device=dev)
    if abc == "PML":   # the fixture's profiles (same formula, vmax of the fixture's vp)
        prop.bcx, prop.bcz = torch.tensor(g["bcx"], device=dev), torch.tensor(g["bcz"], device=dev)
    else:
        prop.damp = torch.tensor(g["damp"], device=dev)
    rec = prop.forward(fd_order=int(g["order"]), checkpoint_segments=int(g["segments"]))
    sum((rec[k] * torch.tensor(g["W_" + k], device=dev)).sum() for k in ("txx", "tzz", "txz")).backward()
    for k in ("vp", "vs", "rho", "eps", "delta"):
        e = rel_l2(getattr(model, k).grad.cpu().numpy(), g[f"g_{k}_stress"])
        assert e <= GRAD_TOL, (name, k, e)


@pytest.mark.parametrize("order", [4, 6])
@pytest.mark.parametrize("shape", [(61, 171), (100, 330)])
def test_fused_abl_large_grid_matches_generic(order, shape):
    """Sponge (ABL) boundary on multi-tile grids: the fused pair ela_f / ela_b against the generic kernels -- ragged edges, receivers on
    tile borders and in the free-surface rows' neighbourhood, duplicate receivers, sources next to tile corners, a sponge plane that
    varies along x in the top rows (so the undamped side buffer of the free-surface rows matters), store-all and checkpointed."""
    from adfwi_b200.propagator import acoustic_kernels as ak, elastic_kernels as ek
    from adfwi_b200.propagator.boundary_condition import bc_gerjan
    dev = torch.device("cuda:0")
    torch.manual_seed(2)
    nz, nx = shape
    nabc, nt, ns = 10, 90, 5
    vp = 2500 + 1000 * torch.rand(nz, nx, device=dev); vs = vp / 1.8; rho = 2000 + 100 * torch.rand(nz, nx, device=dev)
    C33 = vp * vp * rho; C55f = vs * vs * rho; C11 = 1.15 * C33; C13 = C33 - 2 * C55f
    b = 1.0 / rho
    base = dict(C11=C11, C13=C13, C33=C33, C55=C55f[1:-1, 1:-1].clone(), bx=0.5 * (b[:, :-1] + b[:, 1:]), bz=0.5 * (b[:-1] + b[1:]))
    sx = torch.tensor([3, 64, 63, 128, nx - 1], device=dev); sz = torch.tensor([1, 15, 16, 31, 40], device=dev)
    rx = torch.cat([torch.arange(0, nx, 3), torch.tensor([63, 64, 64, 127, 128])]).to(dev)
    rz = torch.cat([torch.full((len(range(0, nx, 3)),), 0), torch.tensor([15, 16, 16, 31, 32])]).to(dev)
    src = torch.randn(ns, nt, device=dev)
    mt = torch.randn(ns, 3, 3, device=dev)
    W = {k: torch.randn(ns, nt, rx.numel(), device=dev) for k in COMPS}
    for fs in (True, False):
        damp = torch.tensor(bc_gerjan(nx, nz, 10.0, 10.0, pml=nabc, alpha=0.02, free_surface=fs), device=dev, dtype=torch.float32)
        damp = damp * (1 - 0.01 * torch.rand_like(damp))         # no symmetry left to hide behind
        out = {}
        for mode in (False, True):
            old = dict(ak.config); ak.config.update(force_generic=mode, shots_per_group=(2 if mode else 0), shots_per_chunk=(0 if fs else 2),
                                                    ckpt_interval=(None if fs else 40))
            try:
                leaves = {k: v.clone().requires_grad_(True) for k, v in base.items()}
                CC = [None] * 21
                CC[0], CC[2], CC[11], CC[18] = leaves["C11"], leaves["C13"], leaves["C33"], leaves["C55"]
                rec = ek.forward_kernel(nx, nz, 10.0, 10.0, nt, 1e-3, nabc, fs, sx, sz, ns, src, mt, rx, rz, rx.numel(), "gerjan", None, None, damp,
                                        None, None, leaves["bx"], leaves["bz"], CC, fd_order=order, n_segments=3, device=dev)
                sum((rec[k] * W[k]).sum() * (1e-6 if k[0] == "t" else 1.0) for k in COMPS).backward()
                out[mode] = (rec, {k: v.grad.clone() for k, v in leaves.items()})
            finally:
                ak.config.clear(); ak.config.update(old)
        for k in COMPS:
            assert torch.equal(out[False][0][k], out[True][0][k]), (order, fs, k)
            assert rel_l2(out[False][0]["forward_wavefield_" + k].cpu().numpy(), out[True][0]["forward_wavefield_" + k].cpu().numpy()) < 1e-5
        for k in PLANES:
            e = rel_l2(out[False][1][k].cpu().numpy(), out[True][1][k].cpu().numpy())
            assert e < 2e-5, (order, fs, k, e)


@pytest.mark.parametrize("order", [4, 6])
@pytest.mark.parametrize("shape", [(61, 171), (100, 330)])
def test_fused_large_grid_matches_generic(order, shape):
    """Multi-tile grid (several 64x16 tiles in x and z, ragged edges, receivers on tile borders, duplicate
    receivers, sources next to tile corners): the TMA-staged split-PML pipeline against the generic kernels.
    The 100x330 grid has all three tile classes of the lean adjoint (PML, damping-free next to PML, damping-free
    with damping-free neighbours: tiles x in {2,3}, z in {0..4})."""
    from adfwi_b200.propagator import acoustic_kernels as ak, elastic_kernels as ek
    from adfwi_b200.propagator.boundary_condition import bc_pml_xz
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    nz, nx = shape
    nabc, nt, ns = 10, 90, 5
    vp = 2500 + 1000 * torch.rand(nz, nx, device=dev); vs = vp / 1.8; rho = 2000 + 100 * torch.rand(nz, nx, device=dev)
    C33 = vp * vp * rho; C55f = vs * vs * rho; C11 = 1.15 * C33; C13 = C33 - 2 * C55f
    b = 1.0 / rho
    base = dict(C11=C11, C13=C13, C33=C33, C55=C55f[1:-1, 1:-1].clone(), bx=0.5 * (b[:, :-1] + b[:, 1:]), bz=0.5 * (b[:-1] + b[1:]))
    sx = torch.tensor([3, 64, 63, 128, nx - 1], device=dev); sz = torch.tensor([1, 15, 16, 31, 40], device=dev)
    rx = torch.cat([torch.arange(0, nx, 3), torch.tensor([63, 64, 64, 127, 128])]).to(dev)
    rz = torch.cat([torch.full((len(range(0, nx, 3)),), 2), torch.tensor([15, 16, 16, 31, 32])]).to(dev)
    src = torch.randn(ns, nt, device=dev)
    mt = torch.randn(ns, 3, 3, device=dev)
    W = {k: torch.randn(ns, nt, rx.numel(), device=dev) for k in COMPS}
    for fs in (True, False):
        bcx, bcz = bc_pml_xz(nx, nz, 10.0, 10.0, pml=nabc, vmax=3500.0, free_surface=fs)
        bcx, bcz = torch.tensor(bcx, device=dev, dtype=torch.float32), torch.tensor(bcz, device=dev, dtype=torch.float32)
        out = {}
        for mode in (False, True):
            old = dict(ak.config); ak.config.update(force_generic=mode, shots_per_group=(2 if mode else 0), shots_per_chunk=(0 if fs else 2),
                                                    ckpt_interval=(None if fs else 40))
            try:
                leaves = {k: v.clone().requires_grad_(True) for k, v in base.items()}
                CC = [None] * 21
                CC[0], CC[2], CC[11], CC[18] = leaves["C11"], leaves["C13"], leaves["C33"], leaves["C55"]
                rec = ek.forward_kernel(nx, nz, 10.0, 10.0, nt, 1e-3, nabc, fs, sx, sz, ns, src, mt, rx, rz, rx.numel(), "PML", bcx, bcz, None,
                                        None, None, leaves["bx"], leaves["bz"], CC, fd_order=order, n_segments=3, device=dev)
                sum((rec[k] * W[k]).sum() * (1e-6 if k[0] == "t" else 1.0) for k in COMPS).backward()
                out[mode] = (rec, {k: v.grad.clone() for k, v in leaves.items()})
            finally:
                ak.config.clear(); ak.config.update(old)
        for k in COMPS:
            assert torch.equal(out[False][0][k], out[True][0][k]), (order, fs, k)
            assert rel_l2(out[False][0]["forward_wavefield_" + k].cpu().numpy(), out[True][0]["forward_wavefield_" + k].cpu().numpy()) < 1e-5
        for k in PLANES:
            e = rel_l2(out[False][1][k].cpu().numpy(), out[True][1][k].cpu().numpy())
            assert e < 2e-5, (order, fs, k, e)


@pytest.mark.parametrize("switch", ["ADFWI_B200_EL_ADJ_SPLIT=1", "ADFWI_B200_EL_LEAN=0", "ADFWI_B200_EL_SPLIT=1", "ADFWI_B200_EL_ADJ_SPLIT=0"])
def test_kernel_variant_switches(golden_dir, switch):
    """The A/B switches of the split-PML pipeline keep working: EL_ADJ_SPLIT=1 = reverse step as the pair elf_k1 + elf_k2
    (the default of O(2,6)), EL_ADJ_SPLIT=0 = the fused reverse kernel also for O(2,6) (elf_b<3>, one CTA per SM), EL_LEAN=0 = full split treatment on damping-free tiles (8 history planes, both halves
    of every pair staged), EL_SPLIT=1 = forward step as elf_s + elf_v.  The switches are read once per process, so the
    golden and large-grid checks run in a child process."""
    import os
    import subprocess
    import sys
    k, v = switch.split("=")
    env = dict(os.environ, **{k: v})
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k",
                        "(test_golden_records_and_gradients and pml_o4 and cfg0) or (fused_large and shape1-4)" if v != "0" else
                        "(test_golden_records_and_gradients and pml_o6 and cfg0) or (fused_large and not abl and shape1-6)"], env=env,
                       capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_moment_tensor_shape_is_validated(golden_dir):
    """MT must be (src_n,3,3); a single (3,3) tensor is broadcast, anything else raises."""
    from adfwi_b200.propagator import elastic_kernels as ek
    g = np.load(f"{golden_dir}/elastic_pml_o4_fs.npz")
    dev = torch.device("cuda:0")
    t = lambda k: torch.tensor(g[k], device=dev)
    nz, nx = int(g["nz"]), int(g["nx"])
    CC = [torch.zeros((nz, nx), device=dev)] * 21
    CC[0], CC[2], CC[11], CC[18] = t("in_C11"), t("in_C13"), t("in_C33"), t("in_C55")
    ns = len(g["src_x"])
    call = lambda mt: ek.forward_kernel(nx, nz, float(g["dx"]), float(g["dz"]), int(g["nt"]), float(g["dt"]), int(g["nabc"]),
                                        bool(g["free_surface"]), t("src_x"), t("src_z"), ns, t("src_v"), mt, t("rcv_x"), t("rcv_z"),
                                        len(g["rcv_x"]), "PML", t("bcx"), t("bcz"), None, None, None, t("in_bx"), t("in_bz"), CC,
                                        fd_order=4, n_segments=1, device=dev)
    with pytest.raises(ValueError, match="MT must have shape"):
        call(t("mt")[: ns - 1])
    one = t("mt")[0]
    a = call(one)
    b = call(one.expand(ns, 3, 3).contiguous())
    for k in COMPS:
        assert torch.equal(a[k], b[k]), k
