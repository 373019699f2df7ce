"""TV / Tikhonov regularisers on the device (``csrc/objective.cu``).

Same class names, constructor and ``forward(m)`` contract as ``ADFWI/fwi/regularization`` (base.py:25-50), including the step
decay of the factors (``regular_StepLR``, base.py:16-18) and the per-call ``iter`` counter -- which ``TV_2order`` does NOT advance
upstream (tv_2order.py:52 returns without ``self.iter += 1``); that quirk is kept.  The value is an autograd scalar; its gradient
with respect to ``m`` is one stencil kernel instead of the autograd walk through nx + nz dense matmuls."""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def regular_StepLR(iter, step_size, alpha, gamma=0.8):
    n = iter // step_size
    return alpha * np.power(gamma, n)


class _Regularizer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, m, kind, nz, nx, dx, dz, alphax, alphaz):
        lib = _lib.load()
        if not m.is_cuda:
            raise RuntimeError("adfwi_b200: the regularisers run only on CUDA tensors (no CPU path)")
        if tuple(m.shape) != (nz, nx):
            raise ValueError(f"adfwi_b200: model plane must be ({nz},{nx})")
        m_c = m.detach().contiguous().float()
        d = _lib.RegularizationDesc()
        d.nz, d.nx, d.kind = int(nz), int(nx), int(kind)
        d.dx, d.dz, d.alphax, d.alphaz = float(dx), float(dz), float(alphax), float(alphaz)
        dev = m_c.device
        with torch.cuda.device(dev):
            wbytes = lib.adfwi_regularization_workspace_bytes(C.byref(d))
            if wbytes == 0:
                raise RuntimeError("adfwi_b200: invalid model dimensions for the regulariser")
            ws = torch.empty(wbytes, dtype=torch.uint8, device=dev)
            val = torch.empty((), dtype=torch.float32, device=dev)
            rc = lib.adfwi_regularization_forward(C.byref(d), m_c.data_ptr(), val.data_ptr(), ws.data_ptr(), wbytes,
                                                  torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib, rc, "adfwi_regularization_forward")
        ctx.held = (d, m_c, ws, wbytes)
        return val

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        d, m_c, ws, wbytes = ctx.held
        dev = m_c.device
        with torch.cuda.device(dev):
            gm = torch.empty_like(m_c)
            gl = g.detach().reshape(1).float().contiguous()
            rc = lib.adfwi_regularization_backward(C.byref(d), m_c.data_ptr(), gl.data_ptr(), gm.data_ptr(), ws.data_ptr(), wbytes,
                                                   torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib, rc, "adfwi_regularization_backward")
        return gm, None, None, None, None, None, None, None


class Regularization:
    kind = None
    advances_iter = True

    def __init__(self, nx, nz, dx, dz, alphax, alphaz, step_size=1000, gamma=1):
        self.iter, self.step_size, self.gamma = 0, step_size, gamma
        self.alphax, self.alphaz = alphax, alphaz
        self.nx, self.nz, self.dx, self.dz = nx, nz, dx, dz

    def forward(self, m: torch.Tensor) -> torch.Tensor:
        ax = regular_StepLR(self.iter, self.step_size, self.alphax, self.gamma)
        az = regular_StepLR(self.iter, self.step_size, self.alphaz, self.gamma)
        out = _Regularizer.apply(m, self.kind, self.nz, self.nx, self.dx, self.dz, ax, az)
        if self.advances_iter:
            self.iter += 1
        return out


class TV_1order(Regularization):
    kind = 0


class Tikhonov_1order(Regularization):
    kind = 1


class TV_2order(Regularization):
    kind = 2
    advances_iter = False          # upstream quirk: tv_2order.py never increments self.iter


class Tikhonov_2order(Regularization):
    kind = 3
