from . import acoustic_kernels, boundary_condition, elastic_kernels  # noqa: F401
from .acoustic_propagator import AcousticPropagator  # noqa: F401
from .boundary_condition import bc_gerjan, bc_pml, bc_pml_xz, bc_sincos  # noqa: F401
from .elastic_propagator import ElasticPropagator  # noqa: F401
from .gradient_process import GradProcessor, smooth2d  # noqa: F401
