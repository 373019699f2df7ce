"""Fused elastic parameterisation (``csrc/parameters.cu``): Thomsen / velocity parameters -> the six coefficient planes of the P-SV
kernels in ONE kernel, and its hand-written transpose in one more -- instead of the ~25 eager elementwise / slicing ops of
``ADFWI/model/parameters.py`` (thomsen_to_elastic_moduli :71-107, elastic_moduli_for_TI :156-181, vs_vp_to_Lame :47-69,
parameter_staggered_grid :184-213) and their autograd mirrors.  Same association and roundings: the planes are bit-identical to the
eager chain, so the records stay bit-identical to the reference."""
import ctypes as C

import torch

from .. import _lib
from ..synthetic import ElasticGridModel

PLANE_KEYS = ("C11", "C13", "C33", "C55", "bx", "bz")


def _shapes(nz, nx):
    return [(nz, nx), (nz, nx), (nz, nx), (nz - 2, nx - 2), (nz, nx - 1), (nz - 1, nx)]


class _ThomsenPlanes(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vp, vs, rho, eps, delta, hti):
        lib = _lib.load()
        ins = [t.detach().contiguous().float() for t in (vp, vs, rho, eps, delta)]
        if not all(t.is_cuda for t in ins):
            raise RuntimeError("adfwi_b200: the fused parameterisation runs only on CUDA tensors (no CPU path)")
        nz, nx = ins[0].shape
        if any(tuple(t.shape) != (nz, nx) for t in ins):
            raise ValueError("adfwi_b200: vp, vs, rho, eps, delta must share one (nz, nx) shape")
        d = _lib.ModuliDesc(); d.nz, d.nx, d.hti = int(nz), int(nx), int(bool(hti))
        dev = ins[0].device
        with torch.cuda.device(dev):
            outs = [torch.empty(s, dtype=torch.float32, device=dev) for s in _shapes(nz, nx)]
            op = _lib.PtrArray6(*[t.data_ptr() for t in outs])
            rc = lib.adfwi_elastic_moduli_forward(C.byref(d), *[t.data_ptr() for t in ins], C.byref(op), torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib, rc, "adfwi_elastic_moduli_forward")
        ctx.held = (d, ins)
        ctx.need = [ctx.needs_input_grad[i] for i in range(5)]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        lib = _lib.load()
        d, ins = ctx.held
        dev = ins[0].device
        nz, nx = ins[0].shape
        with torch.cuda.device(dev):
            gs = [torch.zeros(s, dtype=torch.float32, device=dev) if g is None else g.contiguous().float() for g, s in zip(gs, _shapes(nz, nx))]
            outs = [torch.empty_like(ins[0]) if n else None for n in ctx.need]
            gp = _lib.PtrArray6(*[t.data_ptr() for t in gs])
            rc = lib.adfwi_elastic_moduli_backward(C.byref(d), *[t.data_ptr() for t in ins], C.byref(gp),
                                                   *[None if o is None else o.data_ptr() for o in outs], torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib, rc, "adfwi_elastic_moduli_backward")
        return (*outs, None)


def thomsen_to_staggered_planes(vp, vs, rho, eps, delta, anisotropic_type="vti"):
    """(C11, C13, C33, C55, bx, bz) in the ragged shapes the reference hands to ``forward_kernel``:
    (nz,nx) x3, (nz-2,nx-2), (nz,nx-1), (nz-1,nx).  Isotropic models are the eps = delta = 0 case (the reference's
    IsotropicElasticModel goes through the same ``thomsen_to_elastic_moduli``, elastic_model.py:150-163)."""
    return _ThomsenPlanes.apply(vp, vs, rho, eps, delta, anisotropic_type.lower() == "hti")


class FusedElasticGridModel(ElasticGridModel):
    """``ElasticGridModel`` whose ``forward()`` is the fused kernel: same attributes (``lamu, lam, bx, bz, CC``), two launches per
    forward / backward instead of the eager chain."""

    def __init__(self, *args, anisotropic_type="vti", **kwargs):
        super().__init__(*args, **kwargs)
        self.anisotropic_type = anisotropic_type

    def forward(self):
        C11, C13, C33, C55, bx, bz = thomsen_to_staggered_planes(self.vp, self.vs, self.rho, self.eps, self.delta, self.anisotropic_type)
        zero = torch.zeros_like(self.vp)
        CC = [zero] * 21
        CC[0], CC[2], CC[11], CC[18] = C11, C13, C33, C55
        self.CC, self.bx, self.bz = CC, bx, bz
        self.lamu = self.lam = None          # accepted and unused by forward_kernel (elastic_kernels.py:935-936), as upstream
        return None
