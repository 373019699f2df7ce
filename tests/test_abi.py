"""The C-ABI shared library loads and exports every symbol include/adfwi_b200.h declares
(no compute calls: there is no GPU in the CI container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "adfwi_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(adfwi_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    from adfwi_b200 import _lib
    assert _header_functions() == sorted(_lib.SYMBOLS)


def test_library_builds_loads_and_exports_all_symbols():
    from adfwi_b200 import _lib, build
    path = build.build()
    so = ctypes.CDLL(path)
    for sym in _header_functions():
        assert hasattr(so, sym), f"{sym} missing from libadfwi_b200.so"
    lib = _lib.load()
    assert lib.adfwi_abi_version() == _lib.ABI_VERSION
    assert lib.adfwi_strerror(-3).decode().startswith("adfwi: workspace")
    # argument validation happens before any CUDA call
    d = _lib.AcousticDesc()
    assert lib.adfwi_acoustic_workspace_bytes(ctypes.byref(d)) == 0
    assert lib.adfwi_acoustic_forward(ctypes.byref(d), *([None] * 16), None, 0, None) == -2     # ADFWI_E_DIMS
    e = _lib.ElasticDesc(); e.fd_order = 8
    assert lib.adfwi_elastic_workspace_bytes(ctypes.byref(e)) == 0


def test_struct_layout_matches_header(tmp_path):
    """sizeof / offsetof of every descriptor as gcc lays the header out == the ctypes mirrors of adfwi_b200/_lib.py."""
    import subprocess
    from adfwi_b200 import _lib
    assert ctypes.sizeof(_lib.AcousticDesc) == 4 * 19
    assert ctypes.sizeof(_lib.ElasticDesc) == 4 * 28
    mirrors = {"adfwi_acoustic_desc": _lib.AcousticDesc, "adfwi_elastic_desc": _lib.ElasticDesc, "adfwi_gradproc_desc": _lib.GradProcDesc,
               "adfwi_misfit_desc": _lib.MisfitDesc, "adfwi_regularization_desc": _lib.RegularizationDesc,
               "adfwi_elastic_moduli_desc": _lib.ModuliDesc, "adfwi_elastic_pad_desc": _lib.PadDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "adfwi_b200.h"', 'int main(void) {']
    for cname, cls in mirrors.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0; }")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in mirrors.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


def test_no_cpu_fallback():
    import torch
    from adfwi_b200.propagator import acoustic_kernels as ak
    v = torch.full((20, 24), 2000.0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ak.forward_kernel(24, 20, 10.0, 10.0, 8, 1e-3, 5, True, torch.tensor([3]), torch.tensor([1]), 1,
                          torch.zeros(1, 8), torch.tensor([4]), torch.tensor([1]), 1, torch.zeros(30, 34), v, v.clone(),
                          device="cpu")


def test_missing_library_fails_loudly(monkeypatch):
    from adfwi_b200 import _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libadfwi_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()
