"""Golden fixtures of the gradient post-processing from the UNMODIFIED reference (CPU, float64 numpy/scipy).

    python tests/golden/make_golden_gradproc.py        # needs /root/reference (or $ADFWI_REF)

`scipy.signal.hamming`, which the reference's land taper calls (gradient_process.py:61), no longer exists in the
scipy of this image; it is aliased to the identical `scipy.signal.windows.hamming` before the reference runs.
"""
import os
import sys

import numpy as np
import scipy.signal

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_loader  # noqa: E402

if not hasattr(scipy.signal, "hamming"):
    scipy.signal.hamming = scipy.signal.windows.hamming

CASES = {
    # name: (nz, nx, ctor kwargs, with illumination plane)
    "gradproc_land_full": (52, 70, dict(grad_mute=8, grad_smooth=3, norm_grad=True, forw_illumination=True, marine_or_land="land"), True),
    "gradproc_marine_lower": (52, 70, dict(grad_mute=6, grad_smooth=2, norm_grad=True, forw_illumination=True, marine_or_land="marine"), True),
    "gradproc_marine_caps": (45, 61, dict(grad_mute=5, grad_smooth=0, norm_grad=True, forw_illumination=False, marine_or_land="Marine"), False),
    "gradproc_small_mask": (24, 33, dict(grad_mute=0, grad_smooth=4, norm_grad=False, forw_illumination=True, marine_or_land="Land"), True),
    "gradproc_plain": (30, 41, dict(grad_mute=0, grad_smooth=0, norm_grad=True, forw_illumination=True, marine_or_land="land"), True),
}


def main():
    ref_loader.load()
    from ADFWI.propagator.gradient_process import GradProcessor, smooth2d, grad_taper
    rng = np.random.default_rng(7)
    for name, (nz, nx, kw, with_forw) in CASES.items():
        grad = (rng.standard_normal((nz, nx)) * np.linspace(0.2, 3.0, nz)[:, None]).astype(np.float32)
        forw = (np.abs(rng.standard_normal((nz, nx))) * np.exp(-np.linspace(0, 4, nz))[:, None] * 1e3).astype(np.float32)
        mask = None
        if name.endswith("mask"):
            mask = np.ones((nz, nx)); mask[:, :5] = 0.0; mask[10:14, 20:25] = 0.5
        vmax = np.float32(4123.5)
        gp = GradProcessor(grad_mask=mask, **kw)
        out = gp.forward(nx=nx, nz=nz, vmax=vmax, grad=grad.copy(), forw=forw.copy() if with_forw else None)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), nz=nz, nx=nx, vmax=vmax, grad=grad, forw=forw, with_forw=with_forw,
                            mask=np.zeros(0) if mask is None else mask, out=out, **{"kw_" + k: v for k, v in kw.items()})
        print(name, out.dtype, float(np.abs(out).max()))
    # the two building blocks on their own
    z = rng.standard_normal((37, 55))
    np.savez_compressed(os.path.join(HERE, "gradproc_blocks.npz"), z=z, s3=smooth2d(z, 3), s20=smooth2d(z, 20),
                        taper_land=grad_taper(37, 55, tapersize=9, thred=0.001, marine_or_land="land"),
                        taper_marine=grad_taper(37, 55, tapersize=9, thred=0.0, marine_or_land="Marine"))


if __name__ == "__main__":
    main()
