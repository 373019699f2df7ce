"""GPU parity of the device-side gradient post-processing (adfwi_gradproc_* through the C ABI via the drop-in
GradProcessor / smooth2d) against the reference's golden outputs and against the numpy oracle on a larger plane.
Tolerance: 1e-11 relative to the plane's maximum in float64 (separable vs direct 2-D summation order; the illumination
preconditioner divides by values down to 1e-8, which amplifies nothing relative to the maximum), 1e-6 where the
reference's own dtype flow is float32."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
CASES = ["gradproc_land_full", "gradproc_marine_lower", "gradproc_marine_caps", "gradproc_small_mask", "gradproc_plain"]
TOL = 1e-11


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


def _case(golden_dir, name):
    g = np.load(f"{golden_dir}/{name}.npz")
    kw = {k[3:]: (g[k].item() if g[k].shape == () else g[k]) for k in g.files if k.startswith("kw_")}
    kw["marine_or_land"] = str(kw["marine_or_land"])
    mask = g["mask"] if g["mask"].size else None
    forw = g["forw"] if bool(g["with_forw"]) else None
    return g, kw, mask, forw


def test_smooth2d_blocks(golden_dir):
    from adfwi_b200.propagator import smooth2d
    g = np.load(f"{golden_dir}/gradproc_blocks.npz")
    assert rel(smooth2d(g["z"], 3), g["s3"]) < TOL
    assert rel(smooth2d(g["z"], 20), g["s20"]) < TOL
    zt = torch.tensor(g["z"], device="cuda:0")
    out = smooth2d(zt, 20)
    assert out.is_cuda and rel(out.cpu().numpy(), g["s20"]) < TOL


@pytest.mark.parametrize("name", CASES)
def test_golden(golden_dir, name):
    from adfwi_b200.propagator import GradProcessor
    g, kw, mask, forw = _case(golden_dir, name)
    gp = GradProcessor(grad_mask=mask, **kw)
    out = gp.forward(nx=int(g["nx"]), nz=int(g["nz"]), vmax=g["vmax"][()], grad=g["grad"].copy(), forw=forw)
    assert isinstance(out, np.ndarray) and out.dtype == g["out"].dtype, (out.dtype, g["out"].dtype)
    assert rel(out, g["out"]) < (1e-6 if out.dtype == np.float32 else TOL)
    # device tensors in -> device tensor out, same numbers
    dev = torch.device("cuda:0")
    out_t = gp.forward(nx=int(g["nx"]), nz=int(g["nz"]), vmax=g["vmax"][()], grad=torch.tensor(g["grad"], device=dev),
                       forw=None if forw is None else torch.tensor(forw, device=dev))
    assert out_t.is_cuda and np.array_equal(out_t.cpu().numpy(), out)


@pytest.mark.parametrize("kind,mute,smooth", [("land", 12, 5), ("marine", 9, 4), ("Offshore", 7, 0)])
def test_larger_plane_against_oracle(kind, mute, smooth):
    from adfwi_b200.propagator import GradProcessor
    from oracle import gradproc_oracle as GO
    rng = np.random.default_rng(3)
    nz, nx = 130, 310
    grad = (rng.standard_normal((nz, nx)) * np.linspace(0.1, 2.0, nz)[:, None]).astype(np.float32)
    forw = (np.abs(rng.standard_normal((nz, nx))) * np.exp(-np.linspace(0, 5, nz))[:, None] * 50).astype(np.float32)
    mask = np.ones((nz, nx)); mask[:, -9:] = 0.25
    vmax = np.float32(3999.0)
    ref = GO.grad_process(nx, nz, vmax, grad.copy(), forw=forw, grad_mute=mute, grad_smooth=smooth, grad_mask=mask,
                          norm_grad=True, forw_illumination=True, marine_or_land=kind)
    out = GradProcessor(grad_mute=mute, grad_smooth=smooth, grad_mask=mask, norm_grad=True, forw_illumination=True,
                        marine_or_land=kind).forward(nx=nx, nz=nz, vmax=vmax, grad=grad.copy(), forw=forw)
    assert out.dtype == ref.dtype
    assert rel(out, ref) < TOL


def test_errors():
    from adfwi_b200.propagator import GradProcessor
    with pytest.raises(ValueError):
        GradProcessor(marine_or_land="lunar").forward(nx=8, nz=8, vmax=1.0, grad=np.zeros((8, 8), np.float32))
    with pytest.raises(ValueError):
        GradProcessor(grad_mask=np.ones((3, 3))).forward(nx=8, nz=8, vmax=1.0, grad=np.zeros((8, 8), np.float32))
