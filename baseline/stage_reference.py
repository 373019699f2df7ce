"""Stage the UNMODIFIED reference package for the GPU box (test / benchmark infrastructure only).

    python baseline/stage_reference.py            # copies /root/reference/ADFWI -> baseline/_ref/ADFWI

The reference is pure Python without a build (no setup.py / pyproject: `pip install --target baseline/_ref
/root/reference` has nothing to install), so staging = copying its package directory, .py files only.
``baseline/_ref/`` is git-ignored (the sources never enter this repository's history) but travels with the
`gpurun` snapshot, where /root/reference does not exist.  Consumers: ``oracle/ref_loader.py`` (search order
$ADFWI_REF, /root/reference, baseline/_ref), i.e. tests/test_reference_patch_gpu.py and ``bench.py --impl reference``.
The product package ``adfwi_b200`` never imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def stage(src_root=None, force=False):
    src_root = src_root or os.environ.get("ADFWI_REF") or "/root/reference"
    src = os.path.join(src_root, "ADFWI")
    if not os.path.isdir(os.path.join(src, "propagator")):
        return None
    dst = os.path.join(DST, "ADFWI")
    if os.path.isdir(dst) and not force:
        return DST
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=lambda d, names: [n for n in names if not (n.endswith(".py") or os.path.isdir(os.path.join(d, n)))
                                                        or n == "__pycache__"])
    return DST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv) or "reference tree not found; nothing staged")
