"""ctypes binding of the C ABI declared in include/adfwi_b200.h.

``load()`` returns the product library ``csrc/libadfwi_b200.so`` and raises loudly when it is
missing: there is no CPU / PyTorch fallback anywhere in this package.  ``bind()`` attaches the
prototypes to any CDLL exporting the same ABI (the tests use it for the host-emulation build of
the same sources).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libadfwi_b200.so")
ABI_VERSION = 1

c_f32p = C.c_void_p      # device pointers travel as integers (tensor.data_ptr())
c_i64p = C.c_void_p


class AcousticDesc(C.Structure):
    """adfwi_acoustic_desc (include/adfwi_b200.h)."""
    _fields_ = [
        ("nzp", C.c_int32), ("nxp", C.c_int32),
        ("ns", C.c_int32), ("nt", C.c_int32), ("nr", C.c_int32),
        ("nabc", C.c_int32), ("free_surface", C.c_int32),
        ("dt", C.c_float), ("c1", C.c_float), ("c2", C.c_float),
        ("n_segments", C.c_int32), ("save_history", C.c_int32), ("ckpt_interval", C.c_int32),
        ("need_g_alpha2", C.c_int32), ("shots_per_group", C.c_int32),
        ("reserved", C.c_int32 * 4),
    ]


class ElasticDesc(C.Structure):
    """adfwi_elastic_desc (include/adfwi_b200.h)."""
    _fields_ = [
        ("nzp", C.c_int32), ("nxp", C.c_int32),
        ("ns", C.c_int32), ("nt", C.c_int32), ("nr", C.c_int32),
        ("nz", C.c_int32), ("nx", C.c_int32), ("nabc", C.c_int32),
        ("free_surface", C.c_int32), ("fd_order", C.c_int32), ("abc_pml", C.c_int32),
        ("dt", C.c_float), ("dx", C.c_float), ("dz", C.c_float),
        ("dt_dx", C.c_float), ("dt_dz", C.c_float), ("half_dt", C.c_float),
        ("fdc", C.c_float * 3),
        ("n_segments", C.c_int32), ("save_history", C.c_int32), ("ckpt_interval", C.c_int32),
        ("shots_per_group", C.c_int32),
        ("reserved", C.c_int32 * 4),
    ]


PtrArray5 = C.c_void_p * 5
PtrArray6 = C.c_void_p * 6

# every symbol include/adfwi_b200.h declares (tests/test_abi.py checks the .so exports them all)
SYMBOLS = [
    "adfwi_acoustic_workspace_bytes", "adfwi_acoustic_group_size", "adfwi_acoustic_forward", "adfwi_acoustic_backward",
    "adfwi_elastic_workspace_bytes", "adfwi_elastic_forward", "adfwi_elastic_backward",
    "adfwi_gradproc_workspace_bytes", "adfwi_gradproc_forward", "adfwi_gradproc_smooth2d",
    "adfwi_misfit_workspace_bytes", "adfwi_misfit_forward", "adfwi_misfit_adjoint_source",
    "adfwi_regularization_workspace_bytes", "adfwi_regularization_forward", "adfwi_regularization_backward",
    "adfwi_elastic_moduli_forward", "adfwi_elastic_moduli_backward", "adfwi_elastic_pad_forward", "adfwi_elastic_pad_backward",
    "adfwi_strerror", "adfwi_abi_version", "adfwi_launch_count",
    "adfwi_timing_enable", "adfwi_timing_collect",
]

# enum KernelClass of csrc/common.cuh
KERNEL_CLASSES = [
    "ac_fwd_p", "ac_fwd_uw", "ac_record", "ac_adj_inject", "ac_adj_a", "ac_adj_b",
    "el_fwd_stress", "el_fwd_vel", "el_record", "el_adj_inject", "el_adj_vel", "el_adj_stress",
    "ac_fwd_fused", "ac_adj_fused", "other", "el_fwd_fused", "el_adj_fused", "ac_fwd_persist", "ac_adj_persist",
]


class GradProcDesc(C.Structure):
    """adfwi_gradproc_desc of include/adfwi_b200.h"""
    _fields_ = [("nz", C.c_int32), ("nx", C.c_int32), ("grad_mute", C.c_int32), ("grad_smooth", C.c_int32),
                ("taper_marine", C.c_int32), ("smooth_below_mute", C.c_int32), ("norm_grad", C.c_int32),
                ("use_illumination", C.c_int32), ("illum_span", C.c_int32), ("reserved", C.c_int32),
                ("thred", C.c_double), ("vmax", C.c_double)]


class MisfitDesc(C.Structure):
    """adfwi_misfit_desc of include/adfwi_b200.h"""
    _fields_ = [("ns", C.c_int32), ("nt", C.c_int32), ("nr", C.c_int32), ("kind", C.c_int32), ("normalize", C.c_int32),
                ("reserved", C.c_int32 * 3), ("dt", C.c_double)]


class RegularizationDesc(C.Structure):
    """adfwi_regularization_desc of include/adfwi_b200.h"""
    _fields_ = [("nz", C.c_int32), ("nx", C.c_int32), ("kind", C.c_int32), ("reserved", C.c_int32),
                ("dx", C.c_double), ("dz", C.c_double), ("alphax", C.c_double), ("alphaz", C.c_double)]


class ModuliDesc(C.Structure):
    """adfwi_elastic_moduli_desc of include/adfwi_b200.h"""
    _fields_ = [("nz", C.c_int32), ("nx", C.c_int32), ("hti", C.c_int32), ("reserved", C.c_int32)]


class PadDesc(C.Structure):
    """adfwi_elastic_pad_desc of include/adfwi_b200.h"""
    _fields_ = [("nz", C.c_int32), ("nx", C.c_int32), ("nzp", C.c_int32), ("nxp", C.c_int32), ("pml", C.c_int32), ("top", C.c_int32),
                ("reserved", C.c_int32 * 2)]


def bind(lib):
    vp = C.c_void_p
    if hasattr(lib, "adfwi_elastic_moduli_forward"):
        lib.adfwi_elastic_moduli_forward.restype = C.c_int
        lib.adfwi_elastic_moduli_forward.argtypes = [C.POINTER(ModuliDesc), vp, vp, vp, vp, vp, C.POINTER(PtrArray6), vp]
        lib.adfwi_elastic_moduli_backward.restype = C.c_int
        lib.adfwi_elastic_moduli_backward.argtypes = [C.POINTER(ModuliDesc), vp, vp, vp, vp, vp, C.POINTER(PtrArray6), vp, vp, vp, vp, vp, vp]
        lib.adfwi_elastic_pad_forward.restype = C.c_int
        lib.adfwi_elastic_pad_forward.argtypes = [C.POINTER(PadDesc), C.POINTER(PtrArray6), C.POINTER(PtrArray6), vp]
        lib.adfwi_elastic_pad_backward.restype = C.c_int
        lib.adfwi_elastic_pad_backward.argtypes = [C.POINTER(PadDesc), C.POINTER(PtrArray6), C.POINTER(PtrArray6), vp]
    if hasattr(lib, "adfwi_misfit_forward"):         # absent from the host-emulation fixture of tests/emul
        lib.adfwi_misfit_workspace_bytes.restype = C.c_size_t
        lib.adfwi_misfit_workspace_bytes.argtypes = [C.POINTER(MisfitDesc)]
        lib.adfwi_misfit_forward.restype = C.c_int
        lib.adfwi_misfit_forward.argtypes = [C.POINTER(MisfitDesc), vp, vp, vp, vp, C.c_size_t, vp]
        lib.adfwi_misfit_adjoint_source.restype = C.c_int
        lib.adfwi_misfit_adjoint_source.argtypes = [C.POINTER(MisfitDesc), vp, vp, vp, vp, vp, C.c_size_t, vp]
        lib.adfwi_regularization_workspace_bytes.restype = C.c_size_t
        lib.adfwi_regularization_workspace_bytes.argtypes = [C.POINTER(RegularizationDesc)]
        lib.adfwi_regularization_forward.restype = C.c_int
        lib.adfwi_regularization_forward.argtypes = [C.POINTER(RegularizationDesc), vp, vp, vp, C.c_size_t, vp]
        lib.adfwi_regularization_backward.restype = C.c_int
        lib.adfwi_regularization_backward.argtypes = [C.POINTER(RegularizationDesc), vp, vp, vp, vp, C.c_size_t, vp]
    if hasattr(lib, "adfwi_gradproc_forward"):      # absent from the host-emulation fixture of tests/emul
        lib.adfwi_gradproc_workspace_bytes.restype = C.c_size_t
        lib.adfwi_gradproc_workspace_bytes.argtypes = [C.POINTER(GradProcDesc)]
        lib.adfwi_gradproc_forward.restype = C.c_int
        lib.adfwi_gradproc_forward.argtypes = [C.POINTER(GradProcDesc), vp, vp, vp, vp, C.POINTER(C.c_int), vp, C.c_size_t, vp]
        lib.adfwi_gradproc_smooth2d.restype = C.c_int
        lib.adfwi_gradproc_smooth2d.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_size_t, vp]
    lib.adfwi_acoustic_workspace_bytes.restype = C.c_size_t
    lib.adfwi_acoustic_workspace_bytes.argtypes = [C.POINTER(AcousticDesc)]
    lib.adfwi_acoustic_group_size.restype = C.c_int
    lib.adfwi_acoustic_group_size.argtypes = [C.POINTER(AcousticDesc)]
    lib.adfwi_acoustic_forward.restype = C.c_int
    lib.adfwi_acoustic_forward.argtypes = [C.POINTER(AcousticDesc)] + [vp] * 5 + [vp] * 5 + [vp] * 3 + [vp] * 3 + [vp, C.c_size_t, vp]
    lib.adfwi_acoustic_backward.restype = C.c_int
    lib.adfwi_acoustic_backward.argtypes = [C.POINTER(AcousticDesc)] + [vp] * 5 + [vp] * 5 + [vp] * 3 + [vp] * 3 + [vp, C.c_size_t, vp]
    if hasattr(lib, "adfwi_elastic_forward"):
        lib.adfwi_elastic_workspace_bytes.restype = C.c_size_t
        lib.adfwi_elastic_workspace_bytes.argtypes = [C.POINTER(ElasticDesc)]
        lib.adfwi_elastic_forward.restype = C.c_int
        lib.adfwi_elastic_forward.argtypes = [C.POINTER(ElasticDesc), C.POINTER(PtrArray6), vp, vp, vp, vp, vp, vp, vp, vp,
                                              C.POINTER(PtrArray5), C.POINTER(PtrArray5), vp, C.c_size_t, vp]
        lib.adfwi_elastic_backward.restype = C.c_int
        lib.adfwi_elastic_backward.argtypes = [C.POINTER(ElasticDesc), C.POINTER(PtrArray6), vp, vp, vp, vp, vp, vp, vp, vp,
                                               C.POINTER(PtrArray5), C.POINTER(PtrArray6), vp, vp, C.c_size_t, vp]
    lib.adfwi_strerror.restype = C.c_char_p
    lib.adfwi_strerror.argtypes = [C.c_int]
    lib.adfwi_abi_version.restype = C.c_int
    lib.adfwi_launch_count.restype = C.c_uint64
    lib.adfwi_timing_enable.restype = None
    lib.adfwi_timing_enable.argtypes = [C.c_int]
    lib.adfwi_timing_collect.restype = C.c_int
    lib.adfwi_timing_collect.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int]
    return lib


_LIB = None


def load():
    """The product CUDA library.  Fails loudly (RuntimeError) if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"adfwi_b200: {LIB_PATH} is missing -- build it with `python -m adfwi_b200.build` "
                "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback.")
        lib = bind(C.CDLL(LIB_PATH))
        if lib.adfwi_abi_version() != ABI_VERSION:
            raise RuntimeError("adfwi_b200: libadfwi_b200.so ABI version mismatch; rebuild it")
        _LIB = lib
    return _LIB


def check(lib, rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {lib.adfwi_strerror(rc).decode()} (code {rc})")


def launch_count():
    return int(load().adfwi_launch_count())


def timing_enable(every_n):
    load().adfwi_timing_enable(int(every_n))


def timing_collect():
    """{kernel class: (summed ms of the sampled launches, samples)} and clear the samples."""
    n = len(KERNEL_CLASSES)
    ms = (C.c_float * n)(); cnt = (C.c_int * n)()
    load().adfwi_timing_collect(ms, cnt, n)
    return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(KERNEL_CLASSES) if cnt[i] > 0}
