"""Kernel-LOGIC checks without a GPU: the CUDA sources are compiled for the host by the test
fixture tests/emul (serial emulation of every launch) and compared with the CPU oracle.  This
covers the gather-form adjoints, region masks, free-surface handling, workspace plans,
checkpoint/recompute orchestration and shot grouping.  (The real sm_100a build is exercised by
the -m gpu tests; the product never loads the emulation library.)"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O
import emul_driver as E

COMPS = ("txx", "tzz", "txz", "vx", "vz")
PLANES = ("C11", "C13", "C33", "C55", "bx", "bz")
ELASTIC = [f"elastic_{abc}_o{o}_{fs}" for abc in ("pml", "gerjan") for o in (4, 6) for fs in ("fs", "nofs")]


def _acoustic(g, K, G):
    nabc, fs, dt, dz = int(g["nabc"]), bool(g["free_surface"]), float(g["dt"]), float(g["dz"])
    coef = O.acoustic_coefficients(g["vp"], g["rho"], g["damp"], dt, dz, nabc, fs)
    gr = [g["W_p"], g["W_u"] * 1e7, g["W_w"] * 1e7]
    a = (coef, nabc, fs, dt, g["src_x"], g["src_z"], g["src_v"], g["rcv_x"], g["rcv_z"])
    ref = O.acoustic_run(*a, g_rcv=gr, illum=True, need_g_src=True)
    out = E.acoustic(*a, g_rcv=gr, n_segments=int(g["segments"]), ckpt_interval=K, shots_per_group=G, need_gsrc=True)
    return ref, out


@pytest.mark.parametrize("name", ["acoustic_fs", "acoustic_nofs"])
@pytest.mark.parametrize("K,G", [(0, 0), (40, 1), (64, 2)])
def test_acoustic_kernels_vs_oracle(golden_dir, name, K, G):
    g = np.load(f"{golden_dir}/{name}.npz")
    ref, out = _acoustic(g, K, G)
    for k in "puw":
        assert np.array_equal(out[k], ref[k])
    for k in ("g_alpha1", "g_alpha2", "g_src"):
        assert O.rel_l2(out[k], ref[k]) < 5e-6, k
    for k in "puw":
        assert O.rel_l2(out["illum_" + k], g["rec_forward_wavefield_" + k]) < 1e-5


@pytest.mark.parametrize("name", ELASTIC)
def test_elastic_kernels_vs_oracle(golden_dir, name):
    g = np.load(f"{golden_dir}/{name}.npz")
    planes = {k: g["in_" + k] for k in PLANES}
    kw = dict(bcx=g["bcx"], bcz=g["bcz"]) if str(g["abc"]) == "PML" else dict(damp=g["damp"])
    args = (planes, str(g["abc"]), int(g["order"]), bool(g["free_surface"]), int(g["nz"]), int(g["nx"]), int(g["nabc"]),
            float(g["dx"]), float(g["dz"]), float(g["dt"]), g["src_x"], g["src_z"], g["src_v"], g["mt"], g["rcv_x"], g["rcv_z"])
    gr = [g["W_" + k] * (1e6 if k[0] == "v" else 1.0) for k in COMPS]
    ref = O.elastic_run(*args, g_rcv=gr, n_seg=int(g["segments"]), illum=True, need_g_src=True, **kw)
    K, G = ((0, 0), (30, 1))[ELASTIC.index(name) % 2]
    out = E.elastic(*args, g_rcv=gr, n_seg=int(g["segments"]), ckpt_interval=K, shots_per_group=G, need_gsrc=True, **kw)
    for k in COMPS:
        assert np.array_equal(out[k], ref[k]), k
        assert O.rel_l2(out["illum_" + k], ref["illum_" + k]) < 1e-6
    for k in PLANES:
        assert O.rel_l2(out["g_full"][k], ref["g_full"][k]) < 5e-6, k
    assert O.rel_l2(out["g_src"], ref["g_src"]) < 5e-6


def test_thread_order_independence():
    """Race probe: walking blocks/threads in reverse order must not change a single bit of the
    forward records (an in-place neighbour hazard would)."""
    code = ("import sys, numpy as np; sys.path[:0]=[%r,%r]; import test_emul_kernels as T;"
            "g=np.load(%r); ref,out=T._acoustic(g,40,1);"
            "assert all(np.array_equal(out[k],ref[k]) for k in 'puw');"
            "assert max(T.O.rel_l2(out[k],ref[k]) for k in ('g_alpha1','g_alpha2'))<5e-6")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fix = os.path.join(root, "tests", "golden", "acoustic_fs.npz")
    env = dict(os.environ, ADFWI_EMUL_REVERSE="1")
    subprocess.check_call([sys.executable, "-c", code % (root, os.path.join(root, "tests"), fix)], env=env)
