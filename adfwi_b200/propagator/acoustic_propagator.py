"""``AcousticPropagator`` with the reference's constructor and ``forward()`` contract
(ADFWI/propagator/acoustic_propagator.py:21-157), running on libadfwi_b200.so.

``model`` and ``survey`` are duck-typed: the upstream ``AbstractModel`` / ``Survey`` objects work
unchanged, and so do the light containers of ``adfwi_b200.synthetic``.  What is read from them is
exactly what the reference reads (acoustic_propagator.py:65-98)."""
from typing import Dict, Optional

import numpy as np
import torch
from torch import Tensor

from .acoustic_kernels import forward_kernel, pick_shots
from .boundary_condition import bc_gerjan, bc_pml, bc_sincos

_MODEL_ATTRS = ("ox", "oz", "dx", "dz", "nx", "nz", "abc_type", "nabc", "free_surface", "vp", "rho", "forward")
_SURVEY_ATTRS = ("source", "receiver")


def _to_tensor(a, dtype, device):
    return torch.as_tensor(np.asarray(a), dtype=dtype).to(device)


def _validate(model, survey):
    missing = [a for a in _MODEL_ATTRS if not hasattr(model, a)]
    if missing:
        raise ValueError(f"model is not an instance of AbstractModel (missing {missing})")
    if any(not hasattr(survey, a) for a in _SURVEY_ATTRS):
        raise ValueError("survey is not an instance of Survey")


class AcousticPropagator(torch.nn.Module):
    """Propagator of the isotropic acoustic wave equation (stress-velocity form, staggered-grid FD).

    Parameters mirror the reference: ``model``, ``survey``, ``device``, ``cpu_num``, ``gpu_num``
    (stored, unused -- as upstream), ``dtype`` (float32 only)."""

    def __init__(self, model, survey, device: Optional[str] = "cuda", cpu_num: Optional[int] = 1,
                 gpu_num: Optional[int] = 1, dtype: torch.dtype = torch.float32):
        super().__init__()
        _validate(model, survey)
        self.model, self.survey = model, survey
        self.device, self.dtype = device, dtype
        self.cpu_num, self.gpu_num = cpu_num, gpu_num
        self.ox, self.oz = model.ox, model.oz
        self.dx, self.dz = model.dx, model.dz
        self.nx, self.nz = model.nx, model.nz
        self.nt, self.dt, self.f0 = survey.source.nt, survey.source.dt, survey.source.f0
        self.abc_type, self.nabc, self.free_surface = model.abc_type, model.nabc, model.free_surface
        self.bcx, self.bcz, self.damp = None, None, None
        self.boundary_condition()
        self.source = survey.source
        self.src_loc = self.source.get_loc()
        self.src_x = _to_tensor(self.src_loc[:, 0], torch.long, device)
        self.src_z = _to_tensor(self.src_loc[:, 1], torch.long, device)
        self.src_n = self.source.num
        self.wavelet = _to_tensor(self.source.get_wavelet(), dtype, device)
        self.moment_tensor = _to_tensor(self.source.get_moment_tensor(), dtype, device)
        self.receiver = survey.receiver
        self.rcv_loc = self.receiver.get_loc()
        self.rcv_x = _to_tensor(self.rcv_loc[:, 0], torch.long, device)
        self.rcv_z = _to_tensor(self.rcv_loc[:, 1], torch.long, device)
        self.rcv_n = self.receiver.num

    def boundary_condition(self, vmax=None):
        """Damping plane; always the four-sided profile, vmax frozen at call time
        (acoustic_propagator.py:102-118)."""
        kind = self.abc_type.lower()
        if kind == "pml":
            if vmax is None:
                vmax = self.model.vp.detach().cpu().numpy().max()
            damp = bc_pml(self.nx, self.nz, self.dx, self.dz, pml=self.nabc, vmax=vmax, free_surface=False)
        elif kind == "gerjan":
            damp = bc_gerjan(self.nx, self.nz, self.dx, self.dz, pml=self.nabc,
                             alpha=self.model.abc_jerjan_alpha, free_surface=False)
        else:
            damp = bc_sincos(self.nx, self.nz, self.dx, self.dz, pml=self.nabc, free_surface=False)
        self.damp = _to_tensor(damp, self.dtype, self.device)

    def forward(self, model=None, shot_index=None, checkpoint_segments: int = 1) -> Dict[str, Tensor]:
        """Forward simulation of the selected shots; returns the record dict of ``forward_kernel``."""
        model = self.model if model is None else model
        model.forward()
        pick = lambda t: pick_shots(t, shot_index)
        src_x, src_z, wavelet = pick(self.src_x), pick(self.src_z), pick(self.wavelet)
        return forward_kernel(
            self.nx, self.nz, self.dx, self.dz, self.nt, self.dt,
            self.nabc, self.free_surface,
            src_x, src_z, len(src_x), wavelet,
            self.rcv_x, self.rcv_z, self.rcv_n,
            self.damp, model.vp, model.rho,
            checkpoint_segments=checkpoint_segments, device=self.device, dtype=self.dtype)
