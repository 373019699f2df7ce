"""The numpy oracle of the gradient post-processing against fixtures produced by the unmodified reference
(tests/golden/make_golden_gradproc.py): building blocks and five GradProcessor configurations."""
import numpy as np
import pytest

from oracle import gradproc_oracle as GO

CASES = ["gradproc_land_full", "gradproc_marine_lower", "gradproc_marine_caps", "gradproc_small_mask", "gradproc_plain"]
TOL = 1e-12       # float64 paths; differences are summation order only


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


def load_case(golden_dir, name):
    g = np.load(f"{golden_dir}/{name}.npz")
    kw = {k[3:]: (g[k].item() if g[k].shape == () else g[k]) for k in g.files if k.startswith("kw_")}
    kw["marine_or_land"] = str(kw["marine_or_land"])
    mask = g["mask"] if g["mask"].size else None
    forw = g["forw"] if bool(g["with_forw"]) else None
    return g, kw, mask, forw


def test_blocks(golden_dir):
    g = np.load(f"{golden_dir}/gradproc_blocks.npz")
    assert rel(GO.smooth2d(g["z"], 3), g["s3"]) < TOL
    assert rel(GO.smooth2d(g["z"], 20), g["s20"]) < TOL
    assert rel(GO.taper_plane(37, 55, 9, 0.001, False), g["taper_land"]) < TOL
    assert np.array_equal(GO.taper_plane(37, 55, 9, 0.0, True), g["taper_marine"])


@pytest.mark.parametrize("name", CASES)
def test_grad_process(golden_dir, name):
    g, kw, mask, forw = load_case(golden_dir, name)
    out = GO.grad_process(int(g["nx"]), int(g["nz"]), g["vmax"][()], g["grad"].copy(), forw=forw, grad_mask=mask, **kw)
    assert out.dtype == g["out"].dtype, (out.dtype, g["out"].dtype)
    assert rel(out, g["out"]) < (1e-6 if out.dtype == np.float32 else TOL)
