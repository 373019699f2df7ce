"""Device-side gradient post-processing: drop-in for ``ADFWI.propagator.GradProcessor``
(ADFWI/propagator/gradient_process.py:75-135) and ``smooth2d`` (:30-49).

The reference pulls the gradient to the host, filters it with numpy/scipy in float64 (an 81x81
``convolve2d`` per parameter) and pushes it back (ADFWI/fwi/acoustic_fwi.py:171-177).  Here the same float64
arithmetic runs on the GPU through ``adfwi_gradproc_forward`` (csrc/gradproc.cu), so a gradient that is already a
CUDA tensor never leaves the device.  Same constructor, same ``forward(nx, nz, vmax, grad, forw)``:

* ``grad`` / ``forw`` as CUDA tensors  -> returns a CUDA tensor (float64, or float32 where numpy's promotion
  rules make the reference return float32);
* ``grad`` / ``forw`` as numpy arrays (what the reference's FWI loop passes) -> copied to the current CUDA device,
  processed there, returned as a numpy array of the reference's dtype.

No CPU path: without the CUDA library / a GPU this raises.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def _device_of(*xs):
    for x in xs:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    if not torch.cuda.is_available():
        raise RuntimeError("adfwi_b200.GradProcessor needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _plane(x, dev, dtype, shape):
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
    t = t.detach().to(device=dev, dtype=dtype).contiguous()
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"plane of shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t


def smooth2d(Z, span=10):
    """gradient_process.py:30-49 on the GPU (float64).  numpy in -> numpy out, CUDA tensor in -> CUDA tensor out."""
    lib = _lib.load()
    dev = _device_of(Z)
    z = _plane(Z, dev, torch.float64, np.shape(Z))
    nz, nx = z.shape
    d = _lib.GradProcDesc(nz=nz, nx=nx)
    ws = torch.empty(lib.adfwi_gradproc_workspace_bytes(C.byref(d)), dtype=torch.uint8, device=dev)
    out = torch.empty_like(z)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib, lib.adfwi_gradproc_smooth2d(nz, nx, int(span), z.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), st),
                   "adfwi_gradproc_smooth2d")
    return out if isinstance(Z, torch.Tensor) else out.cpu().numpy()


class GradProcessor():
    def __init__(self, grad_mute=0, grad_smooth=0, grad_mask=None, norm_grad=True, forw_illumination=True, marine_or_land="land"):
        self.grad_mute = grad_mute
        self.grad_smooth = grad_smooth
        self.grad_mask = grad_mask
        self.marine_or_land = marine_or_land
        self.norm_grad = norm_grad
        self.forw_illumination = forw_illumination

    def forward(self, nx, nz, vmax, grad, forw=None):
        kind = self.marine_or_land.lower()
        if kind in ("marine", "offshore"):
            thred = 0.0
        elif kind in ("land", "onshore"):
            thred = 0.001
        else:
            raise ValueError('not supported modeling marine_or_land: %s' % (self.marine_or_land))
        lib = _lib.load()
        as_numpy = not isinstance(grad, torch.Tensor)
        dev = _device_of(grad, forw)
        g = _plane(grad, dev, torch.float32, (nz, nx))
        use_illum = bool(self.forw_illumination) and forw is not None
        f = _plane(forw, dev, torch.float32, (nz, nx)) if use_illum else None
        m = None
        if self.grad_mask is not None:
            if np.shape(self.grad_mask) != (nz, nx):
                raise ValueError('Wrong size of grad mask: the size of the mask should be identical to the size of vp model')
            m = _plane(self.grad_mask, dev, torch.float64, (nz, nx))
        span = 40 if min(nz, nx) > 40 else int(min(nz, nx) / 2)
        d = _lib.GradProcDesc(nz=nz, nx=nx, grad_mute=int(self.grad_mute), grad_smooth=int(self.grad_smooth),
                              taper_marine=int(self.marine_or_land in ('Marine', 'Offshore')),
                              smooth_below_mute=int(self.marine_or_land in ('marine', 'offshore')),
                              norm_grad=int(bool(self.norm_grad)), use_illumination=int(use_illum), illum_span=span,
                              thred=thred, vmax=float(vmax))
        ws = torch.empty(lib.adfwi_gradproc_workspace_bytes(C.byref(d)), dtype=torch.uint8, device=dev)
        out = torch.empty((nz, nx), dtype=torch.float64, device=dev)
        is32 = C.c_int(0)
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream().cuda_stream
            rc = lib.adfwi_gradproc_forward(C.byref(d), g.data_ptr(), f.data_ptr() if f is not None else None,
                                            m.data_ptr() if m is not None else None, out.data_ptr(), C.byref(is32),
                                            ws.data_ptr(), ws.numel(), st)
        _lib.check(lib, rc, "adfwi_gradproc_forward")
        if is32.value:
            out = out.to(torch.float32)          # exact: the values were rounded to float32 on the device
        return out.cpu().numpy() if as_numpy else out
