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
