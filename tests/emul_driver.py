"""TEST FIXTURE: drives the host-emulation build of the CUDA sources (tests/emul) with numpy
buffers through the same C ABI the product uses.  Never imported by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

from adfwi_b200 import _lib
from adfwi_b200.propagator.acoustic_kernels import make_desc

_HERE = os.path.dirname(os.path.abspath(__file__))
_EMUL = None


def emul_lib():
    global _EMUL
    if _EMUL is None:
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "emul"), "-s"])
        _EMUL = _lib.bind(C.CDLL(os.path.join(_HERE, "emul", "libadfwi_emul.so")))
    return _EMUL


def _p(a):
    return None if a is None else a.ctypes.data


def acoustic(coef, nabc, free_surface, dt, src_x, src_z, src_v, rcv_x, rcv_z, g_rcv=None,
             n_segments=1, ckpt_interval=0, shots_per_group=0, need_g2=True, need_gsrc=False):
    lib = emul_lib()
    f32 = np.float32
    planes = [np.ascontiguousarray(coef[k], dtype=f32) for k in ("alpha1", "alpha2", "kappa1", "kappa2", "kappa3")]
    nzp, nxp = planes[0].shape
    src_v = np.ascontiguousarray(src_v, dtype=f32)
    ns, nt = src_v.shape
    nr = len(rcv_x)
    sx = np.ascontiguousarray(np.asarray(src_x) + nabc, dtype=np.int64)
    sz = np.ascontiguousarray(np.asarray(src_z) + nabc, dtype=np.int64)
    rx = np.ascontiguousarray(np.asarray(rcv_x) + nabc, dtype=np.int64)
    rz = np.ascontiguousarray(np.asarray(rcv_z) + nabc, dtype=np.int64)
    save = g_rcv is not None
    d = make_desc(nzp, nxp, ns, nt, nr, nabc, free_surface, dt, n_segments, save, ckpt_interval,
                  need_g2 and save, shots_per_group)
    wb = lib.adfwi_acoustic_workspace_bytes(C.byref(d))
    assert wb > 0
    ws = np.full(wb, 0xFF, dtype=np.uint8)   # poison: the library must initialise what it reads
    rcv = [np.full((ns, nt, nr), np.nan, f32) for _ in range(3)]
    nz, nx = nzp - 2 * nabc, nxp - 2 * nabc
    ill = [np.full((nz, nx), np.nan, f32) for _ in range(3)]
    rc = lib.adfwi_acoustic_forward(C.byref(d), *[_p(a) for a in planes], _p(src_v), _p(sx), _p(sz), _p(rx), _p(rz),
                                    *[_p(a) for a in rcv], *[_p(a) for a in ill], _p(ws), wb, None)
    assert rc == 0, lib.adfwi_strerror(rc)
    out = dict(p=rcv[0], u=rcv[1], w=rcv[2], illum_p=ill[0], illum_u=ill[1], illum_w=ill[2])
    if save:
        gs = [None if g is None else np.ascontiguousarray(g, dtype=f32) for g in g_rcv]
        ga1 = np.full((nzp, nxp), np.nan, f32)
        ga2 = np.full((nzp, nxp), np.nan, f32) if need_g2 else None
        gsrc = np.zeros((ns, nt), f32) if need_gsrc else None
        rc = lib.adfwi_acoustic_backward(C.byref(d), *[_p(a) for a in planes], _p(src_v), _p(sx), _p(sz), _p(rx), _p(rz),
                                         *[_p(a) for a in gs], _p(ga1), _p(ga2), _p(gsrc), _p(ws), wb, None)
        assert rc == 0, lib.adfwi_strerror(rc)
        out.update(g_alpha1=ga1, g_alpha2=ga2, g_src=gsrc)
    return out


def elastic(planes, abc, order, free_surface, nz, nx, nabc, dx, dz, dt, src_x, src_z, src_v, mt, rcv_x, rcv_z,
            bcx=None, bcz=None, damp=None, g_rcv=None, n_seg=1, ckpt_interval=0, shots_per_group=0, need_gsrc=False):
    """Same calling convention as oracle.elastic_run, but through the emulated C ABI."""
    from adfwi_b200.propagator.elastic_kernels import make_desc as el_desc
    from oracle import oracle as O
    lib = emul_lib()
    f32 = np.float32
    NN = 2 if order == 4 else 3
    pml = abc.lower() == "pml"
    nxp = nx + 2 * nabc
    nzp = nz + (nabc + NN if free_surface else 2 * nabc + NN)
    names = ("C11", "C13", "C33", "C55", "bx", "bz")
    full = [O.elastic_full_plane(planes[k], nzp, nxp, nabc, NN, free_surface) for k in names]
    if pml:
        b1 = O.elastic_full_plane(bcx, nzp, nxp, 0, NN, free_surface); b2 = O.elastic_full_plane(bcz, nzp, nxp, 0, NN, free_surface)
    else:
        b1 = O.elastic_full_plane(damp, nzp, nxp, 0, NN, free_surface); b2 = None
    src_v = np.ascontiguousarray(src_v, dtype=f32); mt = np.ascontiguousarray(mt, dtype=f32)
    ns, nt = src_v.shape
    nr = len(rcv_x)
    zoff = NN if free_surface else NN + nabc
    sx = np.ascontiguousarray(np.asarray(src_x) + nabc, dtype=np.int64); sz = np.ascontiguousarray(np.asarray(src_z) + zoff, dtype=np.int64)
    rx = np.ascontiguousarray(np.asarray(rcv_x) + nabc, dtype=np.int64); rz = np.ascontiguousarray(np.asarray(rcv_z) + zoff, dtype=np.int64)
    save = g_rcv is not None
    d = el_desc(nzp, nxp, ns, nt, nr, nz, nx, nabc, free_surface, order, pml, dt, dx, dz, n_seg, save, ckpt_interval, shots_per_group)
    wb = lib.adfwi_elastic_workspace_bytes(C.byref(d))
    assert wb > 0
    ws = np.full(wb, 0xFF, dtype=np.uint8)
    rcv = [np.full((ns, nt, nr), np.nan, f32) for _ in range(5)]
    ill = [np.full((nz, nx), np.nan, f32) for _ in range(5)]
    coef_p = _lib.PtrArray6(*[_p(a) for a in full])
    rcv_p = _lib.PtrArray5(*[_p(a) for a in rcv]); ill_p = _lib.PtrArray5(*[_p(a) for a in ill])
    rc = lib.adfwi_elastic_forward(C.byref(d), C.byref(coef_p), _p(b1), _p(b2), _p(mt), _p(src_v), _p(sx), _p(sz), _p(rx), _p(rz),
                                   C.byref(rcv_p), C.byref(ill_p), _p(ws), wb, None)
    assert rc == 0, lib.adfwi_strerror(rc)
    comps = ("txx", "tzz", "txz", "vx", "vz")
    out = {k: rcv[i] for i, k in enumerate(comps)}
    for i, k in enumerate(comps):
        out["illum_" + k] = ill[i]
    if save:
        gs = [None if g is None else np.ascontiguousarray(g, dtype=f32) for g in g_rcv]
        gc = [np.full((nzp, nxp), np.nan, f32) for _ in range(6)]
        gsrc = np.zeros((ns, nt), f32) if need_gsrc else None
        g_rcv_p = _lib.PtrArray5(*[_p(a) for a in gs]); gc_p = _lib.PtrArray6(*[_p(a) for a in gc])
        rc = lib.adfwi_elastic_backward(C.byref(d), C.byref(coef_p), _p(b1), _p(b2), _p(mt), _p(src_v), _p(sx), _p(sz), _p(rx), _p(rz),
                                        C.byref(g_rcv_p), C.byref(gc_p), _p(gsrc), _p(ws), wb, None)
        assert rc == 0, lib.adfwi_strerror(rc)
        out["g_full"] = dict(zip(names, gc))
        out["g_src"] = gsrc
    return out
