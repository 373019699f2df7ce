"""Dev-time loader for the upstream ADFWI reference (TEST INFRASTRUCTURE ONLY).

Imports the unmodified reference from ``$ADFWI_REF``, ``/root/reference`` or ``baseline/_ref`` with the
plotting / IO third-party modules it does not need on the hot path replaced by
``MagicMock`` stubs.  Used only by ``tests/golden/make_golden.py``
and by the optional ``tests/test_reference_live.py`` (skipped when the reference tree is
absent, e.g. on the GPU box).  Nothing in the product package imports this module.
"""
import os
import sys
from unittest import mock

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1",
    "mpl_toolkits.axes_grid1.axes_divider", "h5py", "obspy", "skimage", "skimage.metrics",
    "ncg_optimizer", "geomloss", "pysdtw", "ot", "ot.lp", "torchinfo", "seaborn", "p_tqdm",
]


_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def reference_root():
    """$ADFWI_REF, /root/reference (the build container), then the copy staged for the GPU box by
    baseline/stage_reference.py (git-ignored, travels with the gpurun snapshot)."""
    for cand in (os.environ.get("ADFWI_REF"), "/root/reference", _STAGED):
        if cand and os.path.isdir(os.path.join(cand, "ADFWI", "propagator")):
            return cand
    return None


def available():
    return reference_root() is not None


def load():
    """Return the imported ``ADFWI`` package of the reference (raises if absent)."""
    root = reference_root()
    if root is None:
        raise ImportError("ADFWI reference tree not found (set $ADFWI_REF)")
    for name in _STUBS:
        try:
            __import__(name)
        except Exception:
            sys.modules.setdefault(name, mock.MagicMock(name=name))
    if root not in sys.path:
        sys.path.insert(0, root)
    import ADFWI  # noqa: F401
    import ADFWI.propagator  # noqa: F401
    import ADFWI.propagator.acoustic_kernels  # noqa: F401
    import ADFWI.propagator.elastic_kernels  # noqa: F401
    return ADFWI
