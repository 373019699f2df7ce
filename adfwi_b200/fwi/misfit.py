"""Fused record post-processing: per-trace max-abs normalisation + misfit + adjoint source, one C-ABI call each way
(``csrc/objective.cu``).

Mirrors the reference's misfit classes -- same names, ``forward(obs, syn)`` argument order and ``dt`` -- with one
addition: ``normalize=True`` folds the reference driver's ``syn / max(|syn|, axis=1, keepdim=True)``
(ADFWI/fwi/acoustic_fwi.py:149-150) into the same two passes over the records, so the ``(ns,nt,nr)`` tensors are read
twice and the adjoint source is written once, instead of the ~10 eager passes (and their autograd mirrors) upstream.
The returned loss is an ordinary autograd scalar: ``loss.backward()`` hands the adjoint source to the propagator."""
import ctypes as C

import torch

from .. import _lib


class _FusedMisfit(torch.autograd.Function):
    @staticmethod
    def forward(ctx, syn, obs, kind, normalize, dt):
        lib = _lib.load()
        if not (syn.is_cuda and obs.is_cuda):
            raise RuntimeError("adfwi_b200: the fused misfit runs only on CUDA tensors (no CPU path)")
        if syn.dim() != 3 or syn.shape != obs.shape:
            raise ValueError("adfwi_b200: syn and obs must both be (n_shots, nt, n_receivers)")
        syn_c, obs_c = syn.detach().contiguous().float(), obs.detach().contiguous().float()
        d = _lib.MisfitDesc()
        d.ns, d.nt, d.nr = (int(v) for v in syn_c.shape)
        d.kind, d.normalize, d.dt = int(kind), int(bool(normalize)), float(dt)
        dev = syn_c.device
        with torch.cuda.device(dev):
            wbytes = lib.adfwi_misfit_workspace_bytes(C.byref(d))
            if wbytes == 0:
                raise RuntimeError("adfwi_b200: invalid record dimensions for the fused misfit")
            ws = torch.empty(wbytes, dtype=torch.uint8, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            rc = lib.adfwi_misfit_forward(C.byref(d), syn_c.data_ptr(), obs_c.data_ptr(), loss.data_ptr(), ws.data_ptr(), wbytes,
                                          torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib, rc, "adfwi_misfit_forward")
        ctx.held = (d, syn_c, obs_c, ws, wbytes)
        return loss

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        d, syn_c, obs_c, ws, wbytes = ctx.held
        dev = syn_c.device
        with torch.cuda.device(dev):
            gs = torch.empty_like(syn_c)
            gl = g.detach().reshape(1).float().contiguous()
            rc = lib.adfwi_misfit_adjoint_source(C.byref(d), syn_c.data_ptr(), obs_c.data_ptr(), gl.data_ptr(), gs.data_ptr(), ws.data_ptr(),
                                                 wbytes, torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib, rc, "adfwi_misfit_adjoint_source")
        return gs, None, None, None, None


class Misfit_waveform_L2:
    """Waveform-difference L2 misfit, ``sum_traces sqrt(sum_t (obs - syn)^2 dt)`` (ADFWI/fwi/misfit/L2.py:12-28)."""
    kind = 0

    def __init__(self, dt=1, normalize=False):
        self.dt, self.normalize = dt, normalize

    def forward(self, obs: torch.Tensor, syn: torch.Tensor) -> torch.Tensor:
        return _FusedMisfit.apply(syn, obs, self.kind, self.normalize, self.dt)

    __call__ = forward


class Misfit_global_correlation(Misfit_waveform_L2):
    """Global-correlation misfit, ``-sum_traces corr(obs/|obs|, syn/|syn|) dt`` (ADFWI/fwi/misfit/GlobalCorrelation.py:13-70)."""
    kind = 1


def normalized_misfit(syn: torch.Tensor, obs: torch.Tensor, misfit=None) -> torch.Tensor:
    """What the reference drivers compute per shot batch with ``waveform_normalize=True``: normalise ``syn`` per trace, then
    ``misfit.forward(obs, syn)`` (acoustic_fwi.py:149-157) -- here in one fused call.  ``obs`` is expected normalised already."""
    misfit = misfit or Misfit_waveform_L2(dt=1)
    return _FusedMisfit.apply(syn, obs, misfit.kind, True, misfit.dt)
