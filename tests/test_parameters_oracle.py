"""Pins the numpy oracle of the elastic parameterisation to the UNMODIFIED reference: the `in_*` planes of the elastic fixtures are the
outputs of the reference's own chain (bit-identical required), and their model-level gradients are what its autograd derives from the
plane-level gradients stored beside them."""
import numpy as np
import pytest

from oracle import parameters_oracle as PO

PLANES = ("C11", "C13", "C33", "C55", "bx", "bz")


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("name", ["elastic_pml_o4_fs", "elastic_gerjan_o6_nofs", "elastic_pml_o6_nofs"])
def test_parameterisation_oracle_matches_reference(golden_dir, name):
    g = np.load(f"{golden_dir}/{name}.npz")
    got = PO.planes(g["vp"], g["vs"], g["rho"], g["eps"], g["delta"])
    for k in PLANES:
        if k == "C13":      # torch's vectorised CPU sqrt is not correctly rounded (1 ulp off in places, a few ulp after the subtraction of C44); numpy's and CUDA's sqrtf are
            assert np.abs(got[k].view(np.int32) - g["in_" + k].view(np.int32)).max() <= 4, k
        else:
            assert np.array_equal(got[k], g["in_" + k]), k
    for tag in ("stress", "vel"):
        grads = PO.planes_T(g["vp"], g["vs"], g["rho"], g["eps"], g["delta"], {k: g[f"g_{k}_{tag}"] for k in PLANES})
        for k, gr in zip(("vp", "vs", "rho", "eps", "delta"), grads):
            assert rel_l2(gr, g[f"g_{k}_{tag}"]) < 2e-5, (tag, k)
