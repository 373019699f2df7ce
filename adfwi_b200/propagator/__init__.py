from . import acoustic_kernels  # noqa: F401
