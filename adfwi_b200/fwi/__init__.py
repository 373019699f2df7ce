"""``adfwi_b200.fwi`` -- what the hot path's immediate callers need on the device:

    fwi.misfit           fused per-trace normalisation + misfit + adjoint source (L2, global correlation)
    fwi.regularization   TV / Tikhonov regularisers of the model planes
    acoustic_gradient / elastic_gradient (this module): the shot-batch loop below

Thin FWI-gradient driver around the propagators: the shot-batch loop of
``AcousticFWI.forward`` (ADFWI/fwi/acoustic_fwi.py:134-166) / ``ElasticFWI.forward``
(ADFWI/fwi/elastic_fwi.py:200-277) reduced to what the hot path needs -- forward, misfit,
backward, gradient accumulation -- so that benchmarks and tests can time "one FWI gradient"
through the public propagator API.  Misfits, regularisers and optimisers stay the upstream
PyTorch ones; ``l2_waveform_misfit`` restates ``Misfit_waveform_L2`` (fwi/misfit/L2.py:22-28)
only so the benchmark is self-contained."""
import math
from typing import Callable, Optional, Sequence

import numpy as np
import torch

from ..propagator.acoustic_kernels import pick_shots


def l2_waveform_misfit(obs: torch.Tensor, syn: torch.Tensor, dt: float = 1.0) -> torch.Tensor:
    """sum over traces of sqrt(sum_t (obs-syn)^2 dt)   (fwi/misfit/L2.py:25-28)."""
    rsd = obs - syn
    return torch.sum(torch.sqrt(torch.sum(rsd * rsd * dt, dim=1)))


def shot_batches(n_shots: int, batch_size: Optional[int]):
    """Contiguous shot batches, last one taking the remainder (acoustic_fwi.py:136-138)."""
    if batch_size is None or batch_size > n_shots:
        batch_size = n_shots
    nb = math.ceil(n_shots / batch_size)
    for b in range(nb):
        lo = b * batch_size
        hi = n_shots if b == nb - 1 else (b + 1) * batch_size
        yield np.arange(lo, hi)


def acoustic_gradient(propagator, obs_p: torch.Tensor, shots: Optional[Sequence[int]] = None,
                      batch_size: Optional[int] = None, misfit: Optional[Callable] = None,
                      checkpoint_segments: int = 1, obs_loader: Optional[Callable] = None):
    """Accumulate d(misfit)/d(model parameters) over ``shots`` into the parameters' ``.grad``.

    ``obs_p`` is indexed by position in ``shots`` ((len(shots), nt, nr)); alternatively
    ``obs_loader(positions) -> tensor`` supplies each batch (used to stream observed data from
    host memory).  Returns (loss tensor on device, summed forward_wavefield_p illumination)."""
    shots = np.arange(propagator.src_n) if shots is None else np.asarray(shots)
    misfit = misfit or (lambda syn, obs: l2_waveform_misfit(obs, syn, propagator.dt))
    total = None
    illum = None
    for pos in shot_batches(len(shots), batch_size):
        rec = propagator.forward(shot_index=shots[pos], checkpoint_segments=checkpoint_segments)
        obs = obs_loader(pos) if obs_loader is not None else pick_shots(obs_p, pos)
        loss = misfit(rec["p"], obs)
        loss.backward()
        total = loss.detach() if total is None else total + loss.detach()
        fw = rec["forward_wavefield_p"]
        illum = fw if illum is None else illum + fw
    return total, illum


def elastic_gradient(propagator, obs: dict, shots: Optional[Sequence[int]] = None, batch_size: Optional[int] = None,
                     components: Sequence[str] = ("vx", "vz"), fd_order: int = 4, checkpoint_segments: int = 1,
                     obs_loader: Optional[Callable] = None):
    """Elastic counterpart of :func:`acoustic_gradient` (the shot-batch loop of ElasticFWI.forward,
    ADFWI/fwi/elastic_fwi.py:200-277): L2 waveform misfit summed over ``components`` of the records.
    ``obs[c]`` is indexed by position in ``shots``; ``obs_loader(positions) -> {c: tensor}`` streams a batch instead.
    Returns (loss tensor on device, summed forward_wavefield_vz illumination)."""
    shots = np.arange(propagator.src_n) if shots is None else np.asarray(shots)
    total = None
    illum = None
    for pos in shot_batches(len(shots), batch_size):
        rec = propagator.forward(shot_index=shots[pos], fd_order=fd_order, checkpoint_segments=checkpoint_segments)
        ob = obs_loader(pos) if obs_loader is not None else {c: pick_shots(obs[c], pos) for c in components}
        loss = sum(l2_waveform_misfit(ob[c], rec[c], propagator.dt) for c in components)
        loss.backward()
        total = loss.detach() if total is None else total + loss.detach()
        fw = rec["forward_wavefield_vz"]
        illum = fw if illum is None else illum + fw
    return total, illum
