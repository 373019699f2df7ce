"""``ElasticPropagator`` with the reference's constructor and ``forward()`` contract
(ADFWI/propagator/elastic_propagator.py:20-142), running on libadfwi_b200.so.  ``model`` /
``survey`` are duck-typed like in :mod:`acoustic_propagator`; the model must expose
``lamu, lam, bx, bz, CC`` after ``model.forward()`` (elastic_propagator.py:138)."""
from typing import Optional

import torch

from .acoustic_kernels import pick_shots
from .acoustic_propagator import _to_tensor
from .boundary_condition import bc_gerjan, bc_pml_xz, bc_sincos
from .elastic_kernels import forward_kernel

_MODEL_ATTRS = ("ox", "oz", "dx", "dz", "nx", "nz", "abc_type", "nabc", "free_surface", "vp", "forward")


class ElasticPropagator(torch.nn.Module):
    """Propagator of the 2-D P-SV elastic wave equation (velocity-stress form, staggered-grid FD)."""

    def __init__(self, model, survey, device: Optional[str] = "cuda", cpu_num: Optional[int] = 1,
                 gpu_num: Optional[int] = 1, dtype=torch.float32):
        super().__init__()
        if any(not hasattr(model, a) for a in _MODEL_ATTRS):
            raise ValueError("model is not AbstractModel")
        if not (hasattr(survey, "source") and hasattr(survey, "receiver")):
            raise ValueError("survey is not Survey")
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

    def boundary_condition(self):
        """Split-PML profiles for the exact spelling "PML" (case-sensitive, as upstream :100),
        Cerjan sponge for 'gerjan', sin/cos sponge otherwise; free-surface aware."""
        vmax = self.model.vp.detach().cpu().numpy().max()
        if self.abc_type == "PML":
            bcx, bcz = bc_pml_xz(self.nx, self.nz, self.dx, self.dz, pml=self.nabc, vmax=vmax, free_surface=self.free_surface)
            self.bcx = _to_tensor(bcx, self.dtype, self.device)
            self.bcz = _to_tensor(bcz, self.dtype, self.device)
        elif self.abc_type == "gerjan":
            damp = bc_gerjan(self.nx, self.nz, self.dx, self.dz, pml=self.nabc, alpha=self.model.abc_jerjan_alpha,
                             free_surface=self.free_surface)
            self.damp = _to_tensor(damp, self.dtype, self.device)
        else:
            damp = bc_sincos(self.nx, self.nz, self.dx, self.dz, pml=self.nabc, free_surface=self.free_surface)
            self.damp = _to_tensor(damp, self.dtype, self.device)

    def forward(self, model=None, shot_index=None, fd_order=4, checkpoint_segments=1):
        model = self.model if model is None else model
        model.forward()
        pick = lambda t: pick_shots(t, shot_index)
        src_x, src_z = pick(self.src_x), pick(self.src_z)
        return forward_kernel(
            self.nx, self.nz, self.dx, self.dz, self.nt, self.dt,
            self.nabc, self.free_surface,
            src_x, src_z, len(src_x), pick(self.wavelet), pick(self.moment_tensor),
            self.rcv_x, self.rcv_z, self.rcv_n,
            self.abc_type, self.bcx, self.bcz, self.damp,
            model.lamu, model.lam, model.bx, model.bz, model.CC,
            fd_order=fd_order, n_segments=checkpoint_segments,
            device=self.device, dtype=self.dtype)
