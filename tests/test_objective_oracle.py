"""Pins the numpy oracle of the record post-processing and the regularisers (oracle/objective_oracle.py) to fixtures of the
UNMODIFIED reference classes (tests/golden/make_golden_objective.py): losses / values to 2e-6 relative, gradients to 2e-5 relative L2
(the reference computes in fp32, the oracle in fp64)."""
import numpy as np
import pytest

from oracle import objective_oracle as OO


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("kind", ["l2", "gc"])
@pytest.mark.parametrize("norm", [0, 1])
@pytest.mark.parametrize("dts", ["1", "s"])
def test_misfit_oracle_matches_reference(golden_dir, kind, norm, dts):
    g = np.load(f"{golden_dir}/objective_misfit.npz")
    dt = 1.0 if dts == "1" else float(g["dt_s"])
    loss, grad = OO.misfit(g["syn"], g["obs"], 0 if kind == "l2" else 1, bool(norm), dt)
    tag = f"{kind}_n{norm}_dt{dts}"
    assert abs(loss - float(g["loss_" + tag])) <= 2e-6 * abs(float(g["loss_" + tag]))
    assert rel_l2(grad, g["g_" + tag]) <= 2e-5, tag


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_regularization_oracle_matches_reference(golden_dir, kind):
    g = np.load(f"{golden_dir}/objective_regularization.npz")
    val, grad = OO.regularization(g["m"], kind, float(g["dx"]), float(g["dz"]), float(g["alphax"]), float(g["alphaz"]))
    assert abs(val - float(g[f"value_{kind}"])) <= 2e-6 * float(g[f"value_{kind}"])
    assert rel_l2(grad, g[f"g_{kind}"]) <= 2e-5, kind
