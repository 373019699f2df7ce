"""Pins the CPU oracle (oracle/) to the golden fixtures produced by the unmodified reference
(tests/golden/make_golden*.py): forward records bit-identical, gradients to 2e-5 relative L2."""
import numpy as np
import pytest

from oracle import oracle as O

COMPS = ("txx", "tzz", "txz", "vx", "vz")
PLANES = ("C11", "C13", "C33", "C55", "bx", "bz")
ELASTIC = [f"elastic_{abc}_o{o}_{fs}" for abc in ("pml", "gerjan") for o in (4, 6) for fs in ("fs", "nofs")]


@pytest.mark.parametrize("name", ["acoustic_fs", "acoustic_nofs"])
def test_acoustic_oracle_matches_reference(golden_dir, name):
    g = np.load(f"{golden_dir}/{name}.npz")
    nabc, fs, dt, dz = int(g["nabc"]), bool(g["free_surface"]), float(g["dt"]), float(g["dz"])
    coef = O.acoustic_coefficients(g["vp"], g["rho"], g["damp"], dt, dz, nabc, fs)
    for comp in "puw":
        gr = [None, None, None]
        gr["puw".index(comp)] = g["W_" + comp]
        out = O.acoustic_run(coef, nabc, fs, dt, g["src_x"], g["src_z"], g["src_v"], g["rcv_x"], g["rcv_z"], g_rcv=gr, illum=True)
        gv, grho = O.acoustic_model_gradients(coef, out["g_alpha1"], out["g_alpha2"], dt, dz, nabc)
        assert O.rel_l2(gv, g["g_v_" + comp]) < 2e-5
        assert O.rel_l2(grho, g["g_rho_" + comp]) < 2e-5
    for k in "puw":
        assert np.array_equal(out[k], g["rec_" + k]), f"record {k} differs from the reference bits"
    nz, nx = int(g["nz"]), int(g["nx"])
    assert O.rel_l2(out["illum_p"][nabc:nabc + nz, nabc:nabc + nx], g["rec_forward_wavefield_p"]) < 1e-5


@pytest.mark.parametrize("name", ELASTIC)
def test_elastic_oracle_matches_reference(golden_dir, name):
    g = np.load(f"{golden_dir}/{name}.npz")
    planes = {k: g["in_" + k] for k in PLANES}
    kw = dict(bcx=g["bcx"], bcz=g["bcz"]) if str(g["abc"]) == "PML" else dict(damp=g["damp"])
    args = (planes, str(g["abc"]), int(g["order"]), bool(g["free_surface"]), int(g["nz"]), int(g["nx"]), int(g["nabc"]),
            float(g["dx"]), float(g["dz"]), float(g["dt"]), g["src_x"], g["src_z"], g["src_v"], g["mt"], g["rcv_x"], g["rcv_z"])
    for tag, use in (("stress", ("txx", "tzz", "txz")), ("vel", ("vx", "vz"))):
        gr = [g["W_" + k] if k in use else None for k in COMPS]
        out = O.elastic_run(*args, g_rcv=gr, n_seg=int(g["segments"]), illum=True, **kw)
        for k in PLANES:
            assert O.rel_l2(out["g_own"][k], g[f"g_{k}_{tag}"]) < 2e-5, (tag, k)
    for k in COMPS:
        assert np.array_equal(out[k], g["rec_" + k]), f"record {k} differs from the reference bits"
        assert O.rel_l2(out["illum_" + k], g["fw_" + k]) < 1e-5


def test_boundary_profiles_match_reference(golden_dir):
    from adfwi_b200.propagator import boundary_condition as bc
    g = np.load(f"{golden_dir}/boundary_profiles.npz")
    for fs in (True, False):
        tag = "fs" if fs else "nofs"
        assert np.array_equal(bc.bc_pml(23, 17, 10.0, 10.0, pml=7, vmax=3210.5, free_surface=fs), g["pml_" + tag])
        bx, bz = bc.bc_pml_xz(23, 17, 10.0, 10.0, pml=7, vmax=3210.5, free_surface=fs)
        assert np.array_equal(bx, g["pmlx_" + tag]) and np.array_equal(bz, g["pmlz_" + tag])
        assert np.array_equal(bc.bc_gerjan(23, 17, 10.0, 10.0, pml=7, alpha=0.0053, free_surface=fs), g["gerjan_" + tag])
        assert np.array_equal(bc.bc_sincos(23, 17, 10.0, 10.0, pml=7, free_surface=fs), g["sincos_" + tag])


def _l2_misfit_and_cotangent(rec, obs, dt):
    """Misfit_waveform_L2 (fwi/misfit/L2.py:22-28) and d(loss)/d(rec) in numpy."""
    r = obs.astype(np.float64) - rec.astype(np.float64)
    n = np.sqrt(np.sum(r * r * dt, axis=1, keepdims=True))
    return float(n.sum()), (-(r * dt) / np.where(n > 0, n, 1.0)).astype(np.float32)


def test_acoustic_oracle_matches_reference_at_example_scale(golden_dir):
    """The oracle against the UNMODIFIED reference at the acoustic example's own size and nt (148 x 260 padded, nt 1600,
    4 shots; tests/golden/make_golden_scale.py): records bit-identical after 1600 steps, loss and vp gradient."""
    g = np.load(f"{golden_dir}/acoustic_c1_scale.npz")
    nabc, dt, dz = int(g["nabc"]), float(g["dt"]), float(g["dz"])
    O.lib().oracle_set_threads(8)
    coef = O.acoustic_coefficients(g["vp_init"], g["rho_init"], g["damp"], dt, dz, nabc, True)
    cot = {}

    def g_rcv(out):
        cot["loss"], gp = _l2_misfit_and_cotangent(out["p"], g["obs_p"], dt)
        return gp, None, None
    out = O.acoustic_run(coef, nabc, True, dt, g["src_x"], g["src_z"], np.broadcast_to(g["wavelet"], (len(g["src_x"]), int(g["nt"]))).copy(),
                         g["rcv_x"], g["rcv_z"], g_rcv=g_rcv, need_g_alpha2=False)
    assert np.array_equal(out["p"], g["rec_p"])
    assert np.array_equal(out["u"][:, :, ::5], g["rec_u"]) and np.array_equal(out["w"][:, :, ::5], g["rec_w"])
    assert abs(cot["loss"] - float(g["loss"])) <= 1e-5 * float(g["loss"])
    gv, _ = O.acoustic_model_gradients(coef, out["g_alpha1"], None, dt, dz, nabc)
    assert O.rel_l2(gv, g["g_vp"]) < 2e-5


def test_elastic_oracle_matches_reference_at_example_scale(golden_dir):
    """The oracle against the UNMODIFIED reference on the VTI example (132 x 280 padded, nt 1000, 2 shots, split-PML O(2,4),
    AnisotropicElasticModel parameterisation): five records bit-identical after 1000 steps, loss, and the eps / delta / vp /
    vs / rho gradients (oracle plane gradients chained through the torch restatement of the parameterisation)."""
    import torch
    from adfwi_b200 import synthetic as syn
    g = np.load(f"{golden_dir}/vti_c4_scale.npz")
    nz, nx, nabc, nt, dt, dx = int(g["nz"]), int(g["nx"]), int(g["nabc"]), int(g["nt"]), float(g["dt"]), float(g["dx"])
    params = ("vp", "vs", "rho", "eps", "delta")
    m = syn.ElasticGridModel(g["vp"], g["vs"], g["rho"], eps=g["eps_init"], delta=g["delta"], dx=dx, dz=dx, nabc=nabc, free_surface=True,
                             abc_type="PML", requires_grad=params, device="cpu")
    m.forward()
    idx = {"C11": 0, "C13": 2, "C33": 11, "C55": 18}
    outs = {k: (m.CC[idx[k]] if k in idx else getattr(m, k)) for k in PLANES}
    planes = {k: v.detach().numpy() for k, v in outs.items()}
    ns = len(g["src_x"])
    O.lib().oracle_set_threads(8)
    cot = {}

    def g_rcv(out):
        lx, gx = _l2_misfit_and_cotangent(out["vx"], g["obs_vx"], dt)
        lz, gz = _l2_misfit_and_cotangent(out["vz"], g["obs_vz"], dt)
        cot["loss"] = lx + lz
        return None, None, None, gx, gz
    out = O.elastic_run(planes, "PML", 4, True, nz, nx, nabc, dx, dx, dt, g["src_x"], g["src_z"], np.broadcast_to(g["wavelet"], (ns, nt)).copy(),
                        np.broadcast_to(np.eye(3, dtype=np.float32), (ns, 3, 3)).copy(), g["rcv_x"], g["rcv_z"], bcx=g["bcx"], bcz=g["bcz"], g_rcv=g_rcv)
    for k, s in (("vx", 1), ("vz", 1), ("txx", 6), ("tzz", 6), ("txz", 6)):
        assert np.array_equal(out[k][:, :, ::s], g["rec_" + k]), f"record {k} differs from the reference bits"
    assert abs(cot["loss"] - float(g["loss"])) <= 1e-5 * float(g["loss"])
    gm = torch.autograd.grad([outs[k] for k in PLANES], [getattr(m, k) for k in params],
                             grad_outputs=[torch.tensor(out["g_own"][k], dtype=torch.float32) for k in PLANES])
    for k, gr in zip(params, gm):
        assert O.rel_l2(gr.numpy(), g["g_" + k]) < 5e-5, k
