"""Drop-in replacement of ``ADFWI/propagator/elastic_kernels.py`` backed by libadfwi_b200.so.

``forward_kernel`` keeps the reference signature and return dict (elastic_kernels.py:917-1037).
The replicate padding of the (ragged) coefficient planes (elastic_kernels.py:176-216, :935-953)
stays in PyTorch so autograd provides its transpose and the Thomsen / Lame chain that produced
``bx, bz, CC`` keeps receiving true autograd edges.  The time loops ``step_forward_PML_*`` /
``step_forward_ABL_*`` and their reverse-mode derivatives run in hand-written sm_100a kernels
behind :class:`ElasticFD`.  There is no CPU path.
"""
import ctypes as C
import struct
from typing import Dict, List

import torch
import torch.nn.functional as F

from .. import _lib
from .acoustic_kernels import _DeferredChecks, _check_indices, _ptr, _require_cuda, config

# fp32 bit patterns of DiffCoef(NN,'s') = inv(A)@B evaluated in fp32 (elastic_kernels.py:20-58):
# O(2,4) 9/8, -1/24; O(2,6) 75/64, -25/384 (one ulp off the nearest float -- an artefact of the
# fp32 matrix inverse that bit-parity requires), 3/640.
_FDC_BITS = {2: (0x3F900000, 0xBD2AAAAB), 3: (0x3F960000, 0xBD855556, 0x3B99999A)}

RECORD_KEYS = ("txx", "tzz", "txz", "vx", "vz")
COEF_KEYS = ("C11", "C13", "C33", "C55", "bx", "bz")


def diff_coef(NN: int):
    return [struct.unpack("<f", struct.pack("<I", b))[0] for b in _FDC_BITS[NN]]


def make_desc(nzp, nxp, ns, nt, nr, nz, nx, nabc, free_surface, fd_order, abc_pml, dt, dx, dz,
              n_segments=1, save_history=False, ckpt_interval=0, shots_per_group=0):
    d = _lib.ElasticDesc()
    d.nzp, d.nxp, d.ns, d.nt, d.nr = int(nzp), int(nxp), int(ns), int(nt), int(nr)
    d.nz, d.nx, d.nabc = int(nz), int(nx), int(nabc)
    d.free_surface, d.fd_order, d.abc_pml = int(bool(free_surface)), int(fd_order), int(bool(abc_pml))
    d.dt, d.dx, d.dz = float(dt), float(dx), float(dz)                 # ctypes rounds to nearest fp32
    d.dt_dx, d.dt_dz, d.half_dt = float(dt) / float(dx), float(dt) / float(dz), 0.5 * float(dt)
    fd = diff_coef(fd_order // 2)
    for k in range(3):
        d.fdc[k] = fd[k] if k < len(fd) else 0.0
    d.n_segments = max(int(n_segments), 1)
    d.save_history = int(bool(save_history))
    d.ckpt_interval = int(ckpt_interval)
    d.shots_per_group = int(shots_per_group)
    return d


def choose_ckpt_interval(lib, desc, budget_bytes):
    nt = desc.nt
    nseg = 1
    while True:
        K = -(-nt // nseg)
        desc.ckpt_interval = 0 if nseg == 1 else K
        need = lib.adfwi_elastic_workspace_bytes(C.byref(desc))
        if need <= budget_bytes or K <= 1:
            return desc.ckpt_interval, need
        nseg += 1


class ElasticFD(torch.autograd.Function):
    """records = time_loop(C11, C13, C33, C55, bx, bz full planes, src_v); hand-written adjoint.

    The six coefficient planes are (nzp,nxp) fp32 (see :func:`full_plane`), bc1/bc2 are the padded
    bcx,bcz (PML) or damp,None (ABL); indices are PADDED grid indices.  Outputs: five records
    (ns,nt,nr) in the order txx,tzz,txz,vx,vz and five illumination maps (nz,nx), non-differentiable."""

    @staticmethod
    def forward(ctx, C11, C13, C33, C55, bx, bz, src_v, bc1, bc2, mt, src_x, src_z, rcv_x, rcv_z,
                nz, nx, nabc, free_surface, fd_order, abc_pml, dt, dx, dz, n_segments):
        lib = _lib.load()
        _require_cuda(C11, C13, C33, C55, bx, bz, src_v, bc1, bc2, mt, src_x, src_z, rcv_x, rcv_z)
        dev = C11.device
        planes = [t.detach().contiguous().float() for t in (C11, C13, C33, C55, bx, bz)]
        src_v_c = src_v.detach().contiguous().float()
        bc1_c = bc1.detach().contiguous().float()
        bc2_c = None if bc2 is None else bc2.detach().contiguous().float()
        mt_c = mt.detach().contiguous().float()
        sx, sz = src_x.contiguous().long(), src_z.contiguous().long()
        rx, rz = rcv_x.contiguous().long(), rcv_z.contiguous().long()
        nzp, nxp = planes[0].shape
        ns, nt = src_v_c.shape
        nr = rx.numel()
        need = [ctx.needs_input_grad[i] for i in range(7)]
        save = any(need)
        desc = make_desc(nzp, nxp, ns, nt, nr, nz, nx, nabc, free_surface, fd_order, abc_pml, dt, dx, dz,
                         n_segments, save, 0, config["shots_per_group"])
        desc.reserved[0] = 1 if config["force_generic"] else 0
        desc.reserved[1] = int(config.get("shots_per_chunk", 0))
        desc.reserved[2] = int(config.get("shots_per_chunk_reverse", 0))
        with torch.cuda.device(dev):
            if save and config["ckpt_interval"] is None:
                free_b, _ = torch.cuda.mem_get_info(dev)
                free_b += torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
                rec_bytes = 5 * ns * nt * max(nr, 1) * 4
                _, wbytes = choose_ckpt_interval(lib, desc, int(config["memory_fraction"] * free_b) - 2 * rec_bytes)
            else:
                if save:
                    desc.ckpt_interval = int(config["ckpt_interval"])
                wbytes = lib.adfwi_elastic_workspace_bytes(C.byref(desc))
            if wbytes == 0:
                raise RuntimeError("adfwi_b200: invalid elastic problem dimensions / fd_order")
            ws = torch.empty(wbytes, dtype=torch.uint8, device=dev)
            rcv = [torch.empty((ns, nt, nr), dtype=torch.float32, device=dev) for _ in range(5)]
            ill = [torch.empty((nz, nx), dtype=torch.float32, device=dev) for _ in range(5)]
            coef_p = _lib.PtrArray6(*[_ptr(t) for t in planes])
            rcv_p = _lib.PtrArray5(*[_ptr(t) for t in rcv])
            ill_p = _lib.PtrArray5(*[_ptr(t) for t in ill])
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = lib.adfwi_elastic_forward(C.byref(desc), C.byref(coef_p), _ptr(bc1_c), _ptr(bc2_c), _ptr(mt_c), _ptr(src_v_c),
                                           _ptr(sx), _ptr(sz), _ptr(rx), _ptr(rz), C.byref(rcv_p), C.byref(ill_p),
                                           _ptr(ws), wbytes, stream)
            _lib.check(lib, rc, "adfwi_elastic_forward")
        ctx.desc, ctx.ws, ctx.wbytes = desc, (ws if save else None), wbytes
        ctx.held = (planes, bc1_c, bc2_c, mt_c, src_v_c, sx, sz, rx, rz)
        ctx.need = need
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(*ill)
        return (*rcv, *ill)

    @staticmethod
    def backward(ctx, g0, g1, g2, g3, g4, *_unused):
        lib = _lib.load()
        planes, bc1_c, bc2_c, mt_c, src_v_c, sx, sz, rx, rz = ctx.held
        desc, ws = ctx.desc, ctx.ws
        if ws is None:
            raise RuntimeError("adfwi_b200: backward called but no history was saved")
        _DeferredChecks.poll()      # input checks of the forward call that have completed since
        dev = planes[0].device
        need = ctx.need
        with torch.cuda.device(dev):
            gs = [None if g is None else g.contiguous().float() for g in (g0, g1, g2, g3, g4)]
            g_coef = [torch.empty_like(planes[0]) for _ in range(6)]
            g_src = torch.zeros_like(src_v_c) if need[6] else None
            coef_p = _lib.PtrArray6(*[_ptr(t) for t in planes])
            g_rcv_p = _lib.PtrArray5(*[_ptr(t) for t in gs])
            g_coef_p = _lib.PtrArray6(*[_ptr(t) for t in g_coef])
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = lib.adfwi_elastic_backward(C.byref(desc), C.byref(coef_p), _ptr(bc1_c), _ptr(bc2_c), _ptr(mt_c), _ptr(src_v_c),
                                            _ptr(sx), _ptr(sz), _ptr(rx), _ptr(rz), C.byref(g_rcv_p), C.byref(g_coef_p),
                                            _ptr(g_src), _ptr(ws), ctx.wbytes, stream)
            _lib.check(lib, rc, "adfwi_elastic_backward")
        ctx.ws = None
        grads = [g_coef[k] if need[k] else None for k in range(6)]
        return (*grads, g_src) + (None,) * 17


def pad_replicate(data: torch.Tensor, pml: int, fs_offset: int, free_surface: bool) -> torch.Tensor:
    """Replicate-pad a plane from ITS OWN shape: ``pml`` columns left/right, ``pml`` rows below,
    ``fs_offset`` (free surface) or ``fs_offset+pml`` rows above -- the values of the reference's
    pad_torchSingle (elastic_kernels.py:176-216), always fp32."""
    top = fs_offset if free_surface else fs_offset + pml
    return F.pad(data.float()[None, None], (pml, pml, top, pml), mode="replicate")[0, 0]


def full_plane(data: torch.Tensor, nzp: int, nxp: int, pml: int, fs_offset: int, free_surface: bool) -> torch.Tensor:
    """The (nzp,nxp) plane the kernels index at region cell (i,j).  The reference pads every
    coefficient from its own ragged shape -- bx (nz,nx-1), bz (nz-1,nx), C55 (nz-2,nx-2),
    ADFWI/model/parameters.py:199-212 -- and then slices it with the full-grid region
    [NN:nzp-NN, NN:nxp-NN] (elastic_kernels.py:303-310); zero-extending the padded plane to the
    full grid reproduces exactly the values that slice sees."""
    p = pad_replicate(data, pml, fs_offset, free_surface)
    h, w = p.shape
    if h > nzp or w > nxp:
        p = p[:nzp, :nxp]
        h, w = p.shape
    if h < nzp or w < nxp:
        p = F.pad(p, (0, nxp - w, 0, nzp - h))
    return p


class _PadPlanes(torch.autograd.Function):
    """The six replicate paddings of forward_kernel (elastic_kernels.py:935-946) + zero extension to the full grid in one kernel, and
    their transpose in the backward pass (``adfwi_elastic_pad_*``).  Takes the planes in the reference's ragged shapes."""

    @staticmethod
    def forward(ctx, C11, C13, C33, C55, bx, bz, nz, nx, nzp, nxp, pml, top):
        lib = _lib.load()
        ins = [t.detach().contiguous().float() for t in (C11, C13, C33, C55, bx, bz)]
        d = _lib.PadDesc(); d.nz, d.nx, d.nzp, d.nxp, d.pml, d.top = int(nz), int(nx), int(nzp), int(nxp), int(pml), int(top)
        dev = ins[0].device
        with torch.cuda.device(dev):
            outs = [torch.empty((nzp, nxp), dtype=torch.float32, device=dev) for _ in range(6)]
            ip, op = _lib.PtrArray6(*[t.data_ptr() for t in ins]), _lib.PtrArray6(*[t.data_ptr() for t in outs])
            rc = lib.adfwi_elastic_pad_forward(C.byref(d), C.byref(ip), C.byref(op), torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib, rc, "adfwi_elastic_pad_forward")
        ctx.d, ctx.shapes = d, [tuple(t.shape) for t in ins]
        ctx.need = [ctx.needs_input_grad[i] for i in range(6)]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        lib = _lib.load()
        dev = next(g.device for g in gs if g is not None)
        with torch.cuda.device(dev):
            gfull = [None if g is None else g.contiguous().float() for g in gs]
            outs = [torch.zeros(s, dtype=torch.float32, device=dev) if (n and gfull[k] is None) else
                    (torch.empty(s, dtype=torch.float32, device=dev) if n else None) for k, (s, n) in enumerate(zip(ctx.shapes, ctx.need))]
            gp = _lib.PtrArray6(*[None if (g is None or o is None) else g.data_ptr() for g, o in zip(gfull, outs)])
            op = _lib.PtrArray6(*[None if (o is None or g is None) else o.data_ptr() for g, o in zip(gfull, outs)])
            rc = lib.adfwi_elastic_pad_backward(C.byref(ctx.d), C.byref(gp), C.byref(op), torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib, rc, "adfwi_elastic_pad_backward")
        return (*outs, None, None, None, None, None, None)


def _ragged_ok(planes, nz, nx):
    want = [(nz, nx), (nz, nx), (nz, nx), (nz - 2, nx - 2), (nz, nx - 1), (nz - 1, nx)]
    return all(torch.is_tensor(t) and t.is_cuda and tuple(t.shape) == w for t, w in zip(planes, want))


def forward_kernel(nx: int, nz: int, dx: float, dz: float, nt: int, dt: float,
                   nabc: int, free_surface: bool,
                   src_x: torch.Tensor, src_z: torch.Tensor, src_n: int, src_v: torch.Tensor, MT: torch.Tensor,
                   rcv_x: torch.Tensor, rcv_z: torch.Tensor, rcv_n: int,
                   abc_type: str, bcx: torch.Tensor, bcz: torch.Tensor, damp: torch.Tensor,
                   lamu: torch.Tensor, lam: torch.Tensor, bx: torch.Tensor, bz: torch.Tensor,
                   CC: List[torch.Tensor],
                   fd_order=4, n_segments=1,
                   device: torch.device = torch.device("cuda"), dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Forward simulation of the elastic wave equation; same contract as the reference's
    ``forward_kernel`` (elastic_kernels.py:917-1037).  ``lamu``/``lam`` are accepted and unused
    (as upstream); ``CC`` is the list of 21 moduli of which C11,C13,C33,C55 (indices 0,2,11,18)
    are used -- C15 and C35 (indices 4,13) are identically zero for every model the package can
    build (ADFWI/model/parameters.py:38-44) and are required to be so here."""
    if dtype != torch.float32:
        raise TypeError("adfwi_b200: only torch.float32 is supported")
    fd_order = 4 if fd_order == 4 else 6          # the reference dispatches `4 else 6` (:981,:997)
    NN = fd_order // 2
    pml = abc_type.lower() in ["pml"]
    C11, C13, C15, C33, C35, C55 = CC[0], CC[2], CC[4], CC[11], CC[13], CC[18]
    _require_cuda(C11, C13, C33, C55, bx, bz, src_v)
    _DeferredChecks.poll()
    tti = [(name, t) for name, t in (("C15", C15), ("C35", C35)) if t is not None and torch.is_tensor(t) and t.is_cuda]
    if tti:       # evaluated on the device, reported without a synchronisation (see _DeferredChecks)
        _DeferredChecks.submit(torch.stack([(t != 0).any() for _, t in tti]),
                               [f"non-zero {name} (TTI) is not supported; the reference never builds it" for name, _ in tti],
                               exc=NotImplementedError)
    if src_v.dim() != 2 or src_v.shape[0] != src_n or src_v.shape[1] != nt:
        raise ValueError("adfwi_b200: src_v must be (src_n, nt)")
    dev = C11.device
    nxp = nx + 2 * nabc
    nzp = nz + (nabc + NN if free_surface else 2 * nabc + NN)
    if config.get("fused_pad", True) and _ragged_ok((C11, C13, C33, C55, bx, bz), nz, nx):
        # the reference's own plane shapes (ADFWI/model/parameters.py:199-212): one fused pad kernel (and one transpose per plane)
        planes = list(_PadPlanes.apply(C11, C13, C33, C55, bx, bz, nz, nx, nzp, nxp, nabc, NN if free_surface else NN + nabc))
    else:
        planes = [full_plane(t, nzp, nxp, nabc, NN, free_surface) for t in (C11, C13, C33, C55, bx, bz)]
    if pml:
        if bcx is None or bcz is None:
            raise ValueError("adfwi_b200: abc_type 'PML' needs bcx and bcz (note: the reference propagator only builds "
                             "them for the exact spelling 'PML', elastic_propagator.py:100)")
        bc1 = full_plane(bcx.to(dev), nzp, nxp, 0, NN, free_surface)
        bc2 = full_plane(bcz.to(dev), nzp, nxp, 0, NN, free_surface)
    else:
        if damp is None:
            raise ValueError("adfwi_b200: sponge boundary needs damp")
        bc1, bc2 = full_plane(damp.to(dev), nzp, nxp, 0, NN, free_surface), None
    src_x, src_z, rcv_x, rcv_z = (t.to(dev) for t in (src_x, src_z, rcv_x, rcv_z))
    _check_indices([("src_x", src_x, nx), ("src_z", src_z, nz), ("rcv_x", rcv_x, nx), ("rcv_z", rcv_z, nz)])
    # the kernels index the moment tensors as mt[s*9 + ...]: (src_n,3,3) is required; a single (3,3) tensor (which the
    # reference's step functions accept for one wavelet) is broadcast
    MT = MT.to(dev).float()
    if MT.dim() == 2 and tuple(MT.shape) == (3, 3):
        MT = MT.expand(src_n, 3, 3)
    if tuple(MT.shape) != (src_n, 3, 3):
        raise ValueError(f"adfwi_b200: MT must have shape ({src_n},3,3) (one moment tensor per selected shot), got {tuple(MT.shape)}")
    zoff = NN if free_surface else NN + nabc
    out = ElasticFD.apply(*planes, src_v.to(dev), bc1, bc2, MT,
                          src_x + nabc, src_z + zoff, rcv_x + nabc, rcv_z + zoff,
                          int(nz), int(nx), int(nabc), bool(free_surface), int(fd_order), bool(pml),
                          float(dt), float(dx), float(dz), int(n_segments))
    rec = {k: out[i] for i, k in enumerate(RECORD_KEYS)}
    for i, k in enumerate(RECORD_KEYS):
        rec["forward_wavefield_" + k] = out[5 + i]
    return rec
