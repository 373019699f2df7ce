"""Model-side helpers on the device: the fused parameterisation of the elastic coefficient planes."""
from .parameters import thomsen_to_staggered_planes, FusedElasticGridModel  # noqa: F401
