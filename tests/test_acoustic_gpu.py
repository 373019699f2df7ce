"""GPU parity of the acoustic path (through the C ABI via the drop-in forward_kernel) against
 (a) the committed golden fixtures of the unmodified reference and (b) the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REC_TOL = 1e-5     # BASELINE.json: synthetic shot records rel-L2 <= 1e-5
GRAD_TOL = 1e-4    # BASELINE.json: gradients rel-L2 <= 1e-4


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _run(g, comp, **cfg):
    from adfwi_b200.propagator import acoustic_kernels as ak
    old = dict(ak.config)
    ak.config.update(cfg)
    try:
        dev = torch.device("cuda:0")
        t = lambda k, **kw: torch.tensor(g[k], device=dev, **kw)
        v = t("vp").requires_grad_(True)
        rho = t("rho").requires_grad_(True)
        rec = ak.forward_kernel(int(g["nx"]), int(g["nz"]), float(g["dx"]), float(g["dz"]), int(g["nt"]), float(g["dt"]),
                                int(g["nabc"]), bool(g["free_surface"]), t("src_x"), t("src_z"), len(g["src_x"]),
                                t("src_v"), t("rcv_x"), t("rcv_z"), len(g["rcv_x"]), t("damp"), v, rho,
                                checkpoint_segments=int(g["segments"]), device=dev, dtype=torch.float32)
        (rec[comp] * t("W_" + comp)).sum().backward()
        return rec, v.grad.cpu().numpy(), rho.grad.cpu().numpy()
    finally:
        ak.config.clear(); ak.config.update(old)


@pytest.mark.parametrize("name", ["acoustic_fs", "acoustic_nofs"])
@pytest.mark.parametrize("cfg", [dict(), dict(ckpt_interval=40, shots_per_group=1), dict(ckpt_interval=64)])
def test_golden_records_and_gradients(golden_dir, name, cfg):
    g = np.load(f"{golden_dir}/{name}.npz")
    for comp in "puw":
        rec, gv, grho = _run(g, comp, **cfg)
        for k in "puw":
            got = rec[k].detach().cpu().numpy()
            assert rel_l2(got, g["rec_" + k]) <= REC_TOL
            # stronger than the bar: the kernels mirror the eager evaluation order bit for bit
            assert np.array_equal(got, g["rec_" + k]), f"{name}: record {k} not bit-identical to the reference"
        assert rel_l2(gv, g[f"g_v_{comp}"]) <= GRAD_TOL, (name, comp)
        assert rel_l2(grho, g[f"g_rho_{comp}"]) <= GRAD_TOL, (name, comp)
        assert rel_l2(gv, g[f"g_v_{comp}"]) <= 2e-5 and rel_l2(grho, g[f"g_rho_{comp}"]) <= 2e-5
    for k in ("forward_wavefield_p", "forward_wavefield_u", "forward_wavefield_w"):
        assert rel_l2(rec[k].cpu().numpy(), g["rec_" + k]) <= 1e-5


def _run_vp_only(g, comp, **cfg):
    from adfwi_b200.propagator import acoustic_kernels as ak
    old = dict(ak.config)
    ak.config.update(cfg)
    try:
        dev = torch.device("cuda:0")
        t = lambda k, **kw: torch.tensor(g[k], device=dev, **kw)
        v = t("vp").requires_grad_(True)
        rec = ak.forward_kernel(int(g["nx"]), int(g["nz"]), float(g["dx"]), float(g["dz"]), int(g["nt"]), float(g["dt"]),
                                int(g["nabc"]), bool(g["free_surface"]), t("src_x"), t("src_z"), len(g["src_x"]),
                                t("src_v"), t("rcv_x"), t("rcv_z"), len(g["rcv_x"]), t("damp"), v, t("rho"),
                                checkpoint_segments=int(g["segments"]), device=dev, dtype=torch.float32)
        (rec[comp] * t("W_" + comp)).sum().backward()
        return rec, v.grad.cpu().numpy()
    finally:
        ak.config.clear(); ak.config.update(old)


@pytest.mark.parametrize("name", ["acoustic_fs", "acoustic_nofs"])
@pytest.mark.parametrize("cfg", [dict(), dict(ckpt_interval=40, shots_per_group=1), dict(ckpt_interval=64),
                                 dict(force_generic=True), dict(shots_per_chunk=2), dict(shots_per_chunk=3, shots_per_group=3),
                                 dict(shots_per_chunk=8, ckpt_interval=50), dict(persistent=False), dict(persistent=False, shots_per_chunk=1)])
def test_fused_pipeline_vp_only(golden_dir, name, cfg):
    """vp-only gradients take the fused TMA pipeline (the default fast path); force_generic
    cross-checks the generic kernels on the same inputs."""
    g = np.load(f"{golden_dir}/{name}.npz")
    for comp in "puw":
        rec, gv = _run_vp_only(g, comp, **cfg)
        for k in "puw":
            got = rec[k].detach().cpu().numpy()
            assert np.array_equal(got, g["rec_" + k]), f"{name}: record {k} not bit-identical to the reference"
        e = rel_l2(gv, g[f"g_v_{comp}"])
        assert e <= GRAD_TOL and e <= 2e-5, (name, comp, e)
    for k in ("forward_wavefield_p", "forward_wavefield_u", "forward_wavefield_w"):
        assert rel_l2(rec[k].cpu().numpy(), g["rec_" + k]) <= 1e-5


def test_fused_large_grid_matches_generic():
    """Multi-tile grid (several 64x32 tiles, ragged edges, many receivers): fused vs generic kernels, vp and rho gradients."""
    from adfwi_b200.propagator import acoustic_kernels as ak
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    nz, nx, nabc, nt, ns = 83, 149, 12, 120, 5
    v = (1800 + 1500 * torch.rand(nz, nx, device=dev))
    rho = 2000 + 200 * torch.rand(nz, nx, device=dev)
    damp = 30 * torch.rand(nz + 2 * nabc, nx + 2 * nabc, device=dev)
    sx = torch.tensor([3, 70, 140, 64, 63], device=dev); sz = torch.tensor([0, 40, 2, 31, 32], device=dev)
    rx = torch.arange(0, nx, 2, device=dev); rz = torch.cat([torch.zeros(40, dtype=torch.long), torch.full((35,), 50)]).to(dev)
    src = torch.randn(ns, nt, device=dev)
    W = torch.randn(ns, nt, rx.numel(), device=dev)
    out = {}
    for fs in (True, False):
        for mode in (False, True):
            old = dict(ak.config); ak.config.update(force_generic=mode, shots_per_group=(2 if mode else 0), shots_per_chunk=(5 if fs else 2))
            try:
                vv = v.clone().requires_grad_(True)
                rr = rho.clone().requires_grad_(True)      # density gradient: fused (SAVE2 / G2 instantiations) vs generic
                rec = ak.forward_kernel(nx, nz, 10.0, 10.0, nt, 1e-3, nabc, fs, sx, sz, ns, src, rx, rz, rx.numel(), damp, vv, rr, device=dev)
                ((rec["p"] * W).sum() + 1e6 * (rec["u"] * W).sum() + 1e6 * (rec["w"] * W).sum()).backward()
                out[mode] = (rec, vv.grad.clone(), rr.grad.clone())
            finally:
                ak.config.clear(); ak.config.update(old)
        for k in ("p", "u", "w"):
            assert torch.equal(out[False][0][k], out[True][0][k]), (fs, k)
        for k in ("forward_wavefield_p", "forward_wavefield_u", "forward_wavefield_w"):
            assert rel_l2(out[False][0][k].cpu().numpy(), out[True][0][k].cpu().numpy()) < 1e-5
        assert rel_l2(out[False][1].cpu().numpy(), out[True][1].cpu().numpy()) < 1e-5, fs
        assert rel_l2(out[False][2].cpu().numpy(), out[True][2].cpu().numpy()) < 1e-5, fs


def test_out_of_range_indices_raise_without_a_per_call_sync(golden_dir):
    """Index validation runs on the device and is reported without a host synchronisation per call
    (acoustic_kernels._DeferredChecks): a bad receiver index raises at the latest at flush_checks()."""
    from adfwi_b200.propagator import acoustic_kernels as ak
    g = np.load(f"{golden_dir}/acoustic_fs.npz")
    dev = torch.device("cuda:0")
    t = lambda k: torch.tensor(g[k], device=dev)
    ak.flush_checks()
    bad_rx = t("rcv_x").clone(); bad_rx[0] = int(g["nx"])
    with pytest.raises(IndexError, match="rcv_x"):
        ak.forward_kernel(int(g["nx"]), int(g["nz"]), float(g["dx"]), float(g["dz"]), int(g["nt"]), float(g["dt"]), int(g["nabc"]),
                          bool(g["free_surface"]), t("src_x"), t("src_z"), len(g["src_x"]), t("src_v"), bad_rx, t("rcv_z"),
                          len(g["rcv_x"]), t("damp"), t("vp"), t("rho"), device=dev)
        ak.flush_checks()
    ak.flush_checks()     # nothing left pending


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_the_same_process(golden_dir):
    """Kernel attributes (the >48 KB dynamic shared-memory opt-in) are per device: forward + backward on cuda:1
    after cuda:0 in one process."""
    from adfwi_b200.propagator import acoustic_kernels as ak, elastic_kernels as ek
    g = np.load(f"{golden_dir}/acoustic_fs.npz")
    for dev in (torch.device("cuda:0"), torch.device("cuda:1")):
        t = lambda k: torch.tensor(g[k], device=dev)
        v = t("vp").requires_grad_(True)
        rec = ak.forward_kernel(int(g["nx"]), int(g["nz"]), float(g["dx"]), float(g["dz"]), int(g["nt"]), float(g["dt"]), int(g["nabc"]),
                                bool(g["free_surface"]), t("src_x"), t("src_z"), len(g["src_x"]), t("src_v"), t("rcv_x"), t("rcv_z"),
                                len(g["rcv_x"]), t("damp"), v, t("rho"), device=dev)
        (rec["p"] * t("W_p")).sum().backward()
        assert np.array_equal(rec["p"].detach().cpu().numpy(), g["rec_p"]), dev
        assert rel_l2(v.grad.cpu().numpy(), g["g_v_p"]) <= 2e-5, dev
    ge = np.load(f"{golden_dir}/elastic_pml_o4_fs.npz")
    for dev in (torch.device("cuda:0"), torch.device("cuda:1")):
        t = lambda k: torch.tensor(ge[k], device=dev)
        nz, nx = int(ge["nz"]), int(ge["nx"])
        L = {k: t("in_" + k).requires_grad_(True) for k in ("C11", "C13", "C33", "C55", "bx", "bz")}
        CC = [torch.zeros((nz, nx), device=dev)] * 21
        CC[0], CC[2], CC[11], CC[18] = L["C11"], L["C13"], L["C33"], L["C55"]
        rec = ek.forward_kernel(nx, nz, float(ge["dx"]), float(ge["dz"]), int(ge["nt"]), float(ge["dt"]), int(ge["nabc"]),
                                bool(ge["free_surface"]), t("src_x"), t("src_z"), len(ge["src_x"]), t("src_v"), t("mt"), t("rcv_x"), t("rcv_z"),
                                len(ge["rcv_x"]), "PML", t("bcx"), t("bcz"), None, None, None, L["bx"], L["bz"], CC, fd_order=4,
                                n_segments=int(ge["segments"]), device=dev)
        sum((rec[k] * t("W_" + k)).sum() for k in ("vx", "vz")).backward()
        assert np.array_equal(rec["vz"].detach().cpu().numpy(), ge["rec_vz"]), dev
        assert rel_l2(L["C11"].grad.cpu().numpy(), ge["g_C11_vel"]) <= 2e-5, dev


@pytest.mark.parametrize("fs", [True, False])
@pytest.mark.parametrize("shape", [(83, 149, 12), (40, 300, 8), (88, 200, 30)])
def test_persistent_small_grid_matches_per_step_kernels(fs, shape):
    """Cluster-persistent kernels (acp_fwd / acp_adj: one launch per sweep, z strips over a cluster, DSMEM halo pulls) against the
    per-step TMA kernels on grids that split into 2..8 strips: records bit-identical, illumination, vp gradient and source gradient."""
    from adfwi_b200.propagator import acoustic_kernels as ak
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    nz, nx, nabc = shape
    nt, ns = 140, 6
    v = (1800 + 1500 * torch.rand(nz, nx, device=dev))
    rho = 2000 + 200 * torch.rand(nz, nx, device=dev)
    damp = 30 * torch.rand(nz + 2 * nabc, nx + 2 * nabc, device=dev)
    sx = torch.tensor([3, nx // 2, nx - 2, 17, 18, nx // 3], device=dev); sz = torch.tensor([0, nz // 2, 2, nz - 1, 1, nz // 3], device=dev)
    rx = torch.cat([torch.arange(0, nx, 2), torch.tensor([5, 5])]).to(dev)
    rz = torch.cat([torch.zeros(len(range(0, nx, 4)), dtype=torch.long), torch.full((len(range(0, nx, 2)) - len(range(0, nx, 4)),), nz // 2),
                    torch.tensor([nz - 1, nz - 1])]).to(dev)
    src = torch.randn(ns, nt, device=dev)
    W = torch.randn(ns, nt, rx.numel(), device=dev)
    out = {}
    for mode in (True, False):
        old = dict(ak.config); ak.config.update(persistent=mode)
        try:
            vv = v.clone().requires_grad_(True)
            ss = src.clone().requires_grad_(True)
            rec = ak.forward_kernel(nx, nz, 10.0, 10.0, nt, 1e-3, nabc, fs, sx, sz, ns, ss, rx, rz, rx.numel(), damp, vv, rho, checkpoint_segments=2, device=dev)
            ((rec["p"] * W).sum() + 1e6 * (rec["u"] * W).sum() + 1e6 * (rec["w"] * W).sum()).backward()
            out[mode] = (rec, vv.grad.clone(), ss.grad.clone())
        finally:
            ak.config.clear(); ak.config.update(old)
    for k in ("p", "u", "w"):
        assert torch.equal(out[True][0][k], out[False][0][k]), (fs, shape, k)
    for k in ("forward_wavefield_p", "forward_wavefield_u", "forward_wavefield_w"):
        assert rel_l2(out[True][0][k].cpu().numpy(), out[False][0][k].cpu().numpy()) < 1e-5, k
    assert rel_l2(out[True][1].cpu().numpy(), out[False][1].cpu().numpy()) < 1e-5
    assert rel_l2(out[True][2].cpu().numpy(), out[False][2].cpu().numpy()) < 1e-5
