"""Builds adfwi_b200/csrc/libadfwi_b200.so with nvcc for sm_100a (in-tree, explicit recipe).

    python -m adfwi_b200.build [--force]

Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   B200 only, no PTX fallback for other parts
  -fmad=false                               one rounding per fp32 op: forward records are
                                            bit-identical to eager PyTorch on CPU
  -lineinfo                                 ncu source pages map to these files
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libadfwi_b200.so")
SOURCES = ["api.cu", "acoustic.cu", "acoustic_fused.cu", "elastic.cu", "elastic_fused.cu", "gradproc.cu", "objective.cu", "parameters.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-Xcompiler", "-fPIC,-O2", "-shared", "-ldl",
    "--threads", "0",          # the eight translation units compile in parallel (elastic_fused.cu alone takes most of the time)
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".inl"))]
    deps.append(os.path.join(HERE, "..", "include", "adfwi_b200.h"))
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)   # the image's CC wrapper is not a usable host compiler for nvcc
    r = subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"], env=env, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libadfwi_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
