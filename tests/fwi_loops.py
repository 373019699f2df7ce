"""Iteration loops used by the iteration-level parity tests (test infrastructure, not product): the reference's
``AcousticFWI.forward`` / ``ElasticFWI.forward`` epoch loops (ADFWI/fwi/acoustic_fwi.py:128-200, elastic_fwi.py:196-320) with
every step on the device -- propagator through the C ABI, fused normalisation + misfit, device-side GradProcessor, torch's own
optimiser and scheduler.  The FWI drivers themselves are out of the hot-path scope."""
from typing import Callable, Optional, Sequence

import numpy as np
import torch

from adfwi_b200.fwi import l2_waveform_misfit, shot_batches


def acoustic_fwi(propagator, model, optimizer, scheduler, obs_p: torch.Tensor, iterations: int, batch_size: Optional[int] = None,
                 gradient_processor=None, waveform_normalize: bool = True, misfit: Optional[Callable] = None,
                 checkpoint_segments: int = 1):
    """The iteration loop of ``AcousticFWI.forward`` (ADFWI/fwi/acoustic_fwi.py:128-200) with every step on the device:
    shot batches -> per-trace max normalisation (:149-150) -> misfit -> backward -> gradient post-processing
    (``adfwi_b200.propagator.GradProcessor``; the reference copies gradient and illumination to the host here,
    :171-177) -> ``optimizer.step()`` / ``scheduler.step()``.  ``optimizer`` / ``scheduler`` are ordinary torch objects over
    ``model.parameters()``; ``misfit(syn, obs)`` defaults to the L2 waveform misfit with dt = 1 as in the examples.
    Returns {"loss": [..], "grad": [processed vp gradient per iteration]} (what the reference caches, :186-189)."""
    misfit = misfit or (lambda syn, obs: l2_waveform_misfit(obs, syn, 1.0))
    obs = obs_p
    if waveform_normalize:
        obs = obs / torch.max(torch.abs(obs), dim=1, keepdim=True).values
    n_shots = propagator.src_n
    hist = {"loss": [], "grad": []}
    for _ in range(iterations):
        optimizer.zero_grad()
        loss_it, forw = 0.0, None
        for pos in shot_batches(n_shots, batch_size):
            rec = propagator.forward(shot_index=pos, checkpoint_segments=checkpoint_segments)
            fw = rec["forward_wavefield_p"]
            forw = fw if forw is None else forw + fw
            syn = rec["p"]
            if waveform_normalize:
                syn = syn / torch.max(torch.abs(syn), dim=1, keepdim=True).values
            loss = misfit(syn, obs[pos])
            loss.backward()
            loss_it += float(loss.item())
        grads = model.vp.grad
        if gradient_processor is not None:
            with torch.no_grad():
                vmax = np.float32(model.vp.detach().max().item())
                grads = gradient_processor.forward(nz=model.nz, nx=model.nx, vmax=vmax, grad=grads, forw=forw)
                model.vp.grad = grads.to(model.vp.dtype)
        optimizer.step()
        scheduler.step()
        hist["loss"].append(loss_it)
        hist["grad"].append(model.vp.grad.detach().clone())
    return hist


def elastic_fwi(propagator, model, optimizer, scheduler, obs: dict, iterations: int, batch_size: Optional[int] = None,
                gradient_processor=None, waveform_normalize: bool = True, components: Sequence[str] = ("vx", "vz"),
                parameters: Sequence[str] = ("vp", "vs", "rho"), fd_order: int = 4, misfit: Optional[Callable] = None,
                checkpoint_segments: int = 1):
    """The iteration loop of ``ElasticFWI.forward`` (ADFWI/fwi/elastic_fwi.py:196-320) for the particle-velocity components
    (``inversion_component`` ``["vx", "vz"]``, for which the reference passes no illumination to the gradient processor,
    :236,247), every step on the device.  ``obs[c]`` are the observed records (ns, nt, nr).  Returns
    {"loss": [..], "grad": {parameter: [processed gradient per iteration]}}."""
    misfit = misfit or (lambda syn, ob: l2_waveform_misfit(ob, syn, 1.0))
    norm = (lambda r: r / torch.max(torch.abs(r), dim=1, keepdim=True).values) if waveform_normalize else (lambda r: r)
    ob = {c: norm(obs[c]) for c in components}
    hist = {"loss": [], "grad": {k: [] for k in parameters}}
    for _ in range(iterations):
        optimizer.zero_grad()
        loss_it = 0.0
        for pos in shot_batches(propagator.src_n, batch_size):
            rec = propagator.forward(shot_index=pos, fd_order=fd_order, checkpoint_segments=checkpoint_segments)
            loss = sum(misfit(norm(rec[c]), ob[c][pos]) for c in components)
            loss.backward()
            loss_it += float(loss.item())
        for k in parameters:
            p = getattr(model, k)
            if gradient_processor is not None and p.grad is not None:
                with torch.no_grad():
                    vmax = np.float32(p.detach().max().item())
                    p.grad = gradient_processor.forward(nz=model.nz, nx=model.nx, vmax=vmax, grad=p.grad, forw=None).to(p.dtype)
            hist["grad"][k].append(p.grad.detach().clone())
        optimizer.step()
        scheduler.step()
        hist["loss"].append(loss_it)
    return hist
