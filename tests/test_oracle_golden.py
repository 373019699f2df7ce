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
