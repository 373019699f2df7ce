"""Drop-in replacement of ``ADFWI/propagator/acoustic_kernels.py`` backed by libadfwi_b200.so.

``forward_kernel`` keeps the reference signature and return dict (acoustic_kernels.py:179-301).
The replicate padding and the coefficient algebra (acoustic_kernels.py:225-265) stay in PyTorch,
so autograd provides their transposes (and whatever produced ``v``/``rho`` -- a leaf Parameter or
a deep-image-prior network -- keeps receiving true autograd edges).  The time loop
(``step_forward``, :41-176) and its reverse-mode derivative run in hand-written sm_100a kernels
behind :class:`AcousticFD`.  There is no CPU path: tensors must live on a CUDA device.
"""
import ctypes as C
import os
from typing import Dict

import torch
import torch.nn.functional as F

from .. import _lib

# FD weights exactly as python evaluates them before they meet an fp32 tensor (:253-254)
C1_STAGGERED = 9.0 / 8.0
C2_STAGGERED = -1.0 / 24.0

# user-tunable knobs (also settable through the environment)
config = {
    # history/checkpoint interval K (time steps); None = pick the largest that fits in memory
    "ckpt_interval": int(os.environ["ADFWI_B200_CKPT_INTERVAL"]) if "ADFWI_B200_CKPT_INTERVAL" in os.environ else None,
    # fraction of the currently free device memory the workspace may take
    "memory_fraction": float(os.environ.get("ADFWI_B200_MEM_FRACTION", "0.85")),
    # shots advanced per kernel launch inside the library (0 = all shots of the call)
    "shots_per_group": int(os.environ.get("ADFWI_B200_SHOTS_PER_GROUP", "0")),
    # shots one CTA of the fused kernels walks through per tile (0 = library picks)
    "shots_per_chunk": int(os.environ.get("ADFWI_B200_SHOTS_PER_CHUNK", "0")),
    # True: run the generic one-cell-per-thread kernels instead of the fused TMA pipeline
    # (cross-checks in the tests; the density gradient always uses the generic kernels)
    "force_generic": os.environ.get("ADFWI_B200_GENERIC", "0") == "1",
    # False: never take the cluster-persistent small-grid kernels (whole time loop in one launch, state resident in shared
    # memory); they are used automatically when the active region fits and the history is store-all
    "persistent": os.environ.get("ADFWI_B200_PERSIST", "1") != "0",
    # elastic shim: pad the six coefficient planes with the fused kernel (False: the eager F.pad chain, kept for cross-checks)
    "fused_pad": os.environ.get("ADFWI_B200_FUSED_PAD", "1") != "0",
}


def make_desc(nzp, nxp, ns, nt, nr, nabc, free_surface, dt, n_segments=1, save_history=False,
              ckpt_interval=0, need_g_alpha2=False, shots_per_group=0):
    d = _lib.AcousticDesc()
    d.nzp, d.nxp, d.ns, d.nt, d.nr = int(nzp), int(nxp), int(ns), int(nt), int(nr)
    d.nabc, d.free_surface = int(nabc), int(bool(free_surface))
    d.dt, d.c1, d.c2 = float(dt), C1_STAGGERED, C2_STAGGERED      # ctypes rounds to fp32 (nearest)
    d.n_segments = max(int(n_segments), 1)
    d.save_history = int(bool(save_history))
    d.ckpt_interval = int(ckpt_interval)
    d.need_g_alpha2 = int(bool(need_g_alpha2))
    d.shots_per_group = int(shots_per_group)
    return d


def choose_ckpt_interval(lib, desc, budget_bytes):
    """Largest history interval K (fewest recomputed steps) whose workspace fits the budget."""
    nt = desc.nt
    nseg = 1
    while True:
        K = -(-nt // nseg)
        desc.ckpt_interval = 0 if nseg == 1 else K
        need = lib.adfwi_acoustic_workspace_bytes(C.byref(desc))
        if need <= budget_bytes or K <= 1:
            return desc.ckpt_interval, need
        nseg += 1


def _ptr(t):
    return None if t is None else t.data_ptr()


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("adfwi_b200: tensors must be on a CUDA device -- this package has no CPU path "
                               "(the wave propagation runs only in libadfwi_b200.so, sm_100a)")


class AcousticFD(torch.autograd.Function):
    """records = time_loop(alpha1, alpha2, kappa1, kappa2, kappa3, src_v); hand-written adjoint.

    Inputs are the padded coefficient planes of acoustic_kernels.py:257-265 and PADDED grid
    indices.  Outputs: rcv_p, rcv_u, rcv_w (ns,nt,nr) and the three illumination maps (nz,nx)
    (non-differentiable, like the reference's ``.detach()``-ed forward wavefields).
    """

    @staticmethod
    def forward(ctx, alpha1, alpha2, kappa1, kappa2, kappa3, src_v, src_x, src_z, rcv_x, rcv_z,
                nabc, free_surface, dt, n_segments):
        lib = _lib.load()
        _require_cuda(alpha1, alpha2, kappa1, kappa2, kappa3, src_v, src_x, src_z, rcv_x, rcv_z)
        dev = alpha1.device
        planes = [t.detach().contiguous().float() for t in (alpha1, alpha2, kappa1, kappa2, kappa3)]
        src_v_c = src_v.detach().contiguous().float()
        sx, sz = src_x.contiguous().long(), src_z.contiguous().long()
        rx, rz = rcv_x.contiguous().long(), rcv_z.contiguous().long()
        nzp, nxp = planes[0].shape
        ns, nt = src_v_c.shape
        nr = rx.numel()
        nz, nx = nzp - 2 * nabc, nxp - 2 * nabc
        need = [ctx.needs_input_grad[i] for i in range(6)]
        save = need[0] or need[1] or need[5]
        desc = make_desc(nzp, nxp, ns, nt, nr, nabc, free_surface, dt, n_segments, save,
                         0, need[1], config["shots_per_group"])
        desc.reserved[0] = (1 if config["force_generic"] else 0) | (0 if config.get("persistent", True) else 2)
        desc.reserved[1] = int(config.get("shots_per_chunk", 0))
        desc.reserved[2] = int(config.get("shots_per_chunk_reverse", 0))
        with torch.cuda.device(dev):
            if save:
                if config["ckpt_interval"] is not None:
                    desc.ckpt_interval = int(config["ckpt_interval"])
                    wbytes = lib.adfwi_acoustic_workspace_bytes(C.byref(desc))
                else:
                    free_b, _ = torch.cuda.mem_get_info(dev)
                    free_b += torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
                    rec_bytes = 3 * ns * nt * max(nr, 1) * 4
                    budget = int(config["memory_fraction"] * free_b) - 2 * rec_bytes
                    _, wbytes = choose_ckpt_interval(lib, desc, budget)
            else:
                wbytes = lib.adfwi_acoustic_workspace_bytes(C.byref(desc))
            if wbytes == 0:
                raise RuntimeError("adfwi_b200: invalid acoustic problem dimensions")
            ws = torch.empty(wbytes, dtype=torch.uint8, device=dev)
            rcv = [torch.empty((ns, nt, nr), dtype=torch.float32, device=dev) for _ in range(3)]
            ill = [torch.empty((nz, nx), dtype=torch.float32, device=dev) for _ in range(3)]
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = lib.adfwi_acoustic_forward(C.byref(desc), *[_ptr(t) for t in planes], _ptr(src_v_c), _ptr(sx), _ptr(sz),
                                            _ptr(rx), _ptr(rz), *[_ptr(t) for t in rcv], *[_ptr(t) for t in ill],
                                            _ptr(ws), wbytes, stream)
            _lib.check(lib, rc, "adfwi_acoustic_forward")
        ctx.desc, ctx.ws, ctx.wbytes = desc, (ws if save else None), wbytes
        ctx.held = (planes, src_v_c, sx, sz, rx, rz)
        ctx.need = need
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(*ill)
        return (*rcv, *ill)

    @staticmethod
    def backward(ctx, g_p, g_u, g_w, *_unused):
        lib = _lib.load()
        planes, src_v_c, sx, sz, rx, rz = ctx.held
        desc, ws = ctx.desc, ctx.ws
        if ws is None:
            raise RuntimeError("adfwi_b200: backward called but no history was saved")
        _DeferredChecks.poll()      # input checks of the forward call that have completed since
        dev = planes[0].device
        need = ctx.need
        with torch.cuda.device(dev):
            gs = [None if g is None else g.contiguous().float() for g in (g_p, g_u, g_w)]
            g_a1 = torch.empty_like(planes[0])
            g_a2 = torch.empty_like(planes[0]) if need[1] else None
            g_src = torch.zeros_like(src_v_c) if need[5] else None
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = lib.adfwi_acoustic_backward(C.byref(desc), *[_ptr(t) for t in planes], _ptr(src_v_c), _ptr(sx), _ptr(sz),
                                             _ptr(rx), _ptr(rz), *[_ptr(t) for t in gs], _ptr(g_a1), _ptr(g_a2), _ptr(g_src),
                                             _ptr(ws), ctx.wbytes, stream)
            _lib.check(lib, rc, "adfwi_acoustic_backward")
        ctx.ws = None   # release the history
        return (g_a1 if need[0] else None, g_a2, None, None, None, g_src,
                None, None, None, None, None, None, None, None)


def pad_replicate(v: torch.Tensor, pml: int) -> torch.Tensor:
    """(nz,nx) -> (nz+2*pml, nx+2*pml), edge values replicated; same values as the reference's
    pad_torchSingle (acoustic_kernels.py:17-37), always fp32."""
    return F.pad(v.float()[None, None], (pml, pml, pml, pml), mode="replicate")[0, 0]


_SCALARS = {}


def _device_scalar(value, device):
    """0-dim fp32 device tensor, created once per (value, device): building it per call is a blocking host-to-device copy, i.e. a
    stream synchronisation on every forward."""
    key = (value, str(device))
    t = _SCALARS.get(key)
    if t is None:
        t = _SCALARS[key] = torch.tensor(value, dtype=torch.float32, device=device)
    return t


def pick_shots(t, shot_index):
    """``t[shot_index]`` as the reference propagators write it (acoustic_propagator.py:142-145); a contiguous ascending index -- what the
    reference's shot batches are (acoustic_fwi.py:136-138) -- becomes a slice: no index tensor, no host-to-device copy, no sync."""
    if shot_index is None:
        return t
    if isinstance(shot_index, slice):
        return t[shot_index]
    import numpy as np
    idx = np.asarray(shot_index.cpu() if torch.is_tensor(shot_index) else shot_index)
    if idx.ndim == 1 and idx.size and idx.dtype.kind in "iu" and idx[0] >= 0 and np.array_equal(idx, np.arange(idx[0], idx[0] + idx.size)):
        return t[int(idx[0]):int(idx[0]) + idx.size]
    return t[shot_index]


def coefficient_planes(v, rho, damp, dt, dz, nabc, free_surface):
    """alpha1, alpha2, kappa1, kappa2, kappa3 with the reference's rounding order (:257-265)."""
    c = pad_replicate(v, nabc)
    den = pad_replicate(rho, nabc)
    nzp, nxp = c.shape
    fs = nabc if free_surface else 1
    # `tensor / python_float` on CUDA multiplies by the rounded reciprocal; the CPU reference (the
    # parity oracle) performs a true division, so divide by a 0-dim device tensor instead.
    dz_t = _device_scalar(float(dz), c.device)
    alpha1 = den * c * c * dt / dz_t
    kappa1 = damp * dt
    alpha2 = dt / (den * dz)
    kappa2 = torch.zeros_like(damp)
    kappa2[:, 1:nxp - 2] = 0.5 * (damp[:, 1:nxp - 2] + damp[:, 2:nxp - 1]) * dt
    kappa3 = torch.zeros_like(damp)
    kappa3[fs:nzp - 2, :] = 0.5 * (damp[fs:nzp - 2, :] + damp[fs + 1:nzp - 1, :]) * dt
    return alpha1, alpha2, kappa1, kappa2, kappa3


class _DeferredChecks:
    """Input validation without a host-device synchronisation per call.

    The range test of the source / receiver indices (and the zero test of the TTI moduli in the elastic
    shim) is evaluated ON THE DEVICE; the resulting flags travel to pinned host memory with a
    non-blocking copy and are read once their event has completed -- at the start of the next
    ``forward_kernel`` call, in ``backward`` or through :func:`flush_checks`.  A bad input therefore
    raises one call late, but it still raises (the kernels themselves skip out-of-grid cells, so
    nothing is read or written out of bounds in the meantime).  ``ADFWI_B200_SYNC_CHECKS=1`` makes
    every check blocking."""
    pending = []
    blocking = os.environ.get("ADFWI_B200_SYNC_CHECKS", "0") == "1"

    @classmethod
    def submit(cls, flags, messages, exc=IndexError):
        host = torch.empty(flags.shape, dtype=flags.dtype, pin_memory=True)
        host.copy_(flags, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(flags.device))
        cls.pending.append((ev, host, messages, exc))
        cls.poll(block=cls.blocking)

    @classmethod
    def poll(cls, block=False):
        keep, pend = [], cls.pending
        cls.pending = []
        err = None
        for item in pend:
            ev, host, messages, exc = item
            if block:
                ev.synchronize()
            if not ev.query():
                keep.append(item)
                continue
            bad = [m for m, f in zip(messages, host.tolist()) if f]
            if bad and err is None:
                err = exc("adfwi_b200: " + "; ".join(bad))
        cls.pending = keep + cls.pending
        if err is not None:
            raise err


def flush_checks():
    """Block until every deferred input check has been evaluated; raises if one failed."""
    _DeferredChecks.poll(block=True)


def _check_indices(pairs):
    """pairs: [(name, index tensor on the device, exclusive upper bound)]; see :class:`_DeferredChecks`."""
    pairs = [(n, t, hi) for n, t, hi in pairs if t.numel()]
    if not pairs:
        return
    flags = torch.stack([(t.min() < 0) | (t.max() >= hi) for _, t, hi in pairs])
    _DeferredChecks.submit(flags, [f"{n} out of range for a grid of {hi} points" for n, _, hi in pairs])


def forward_kernel(nx: int, nz: int, dx: float, dz: float, nt: int, dt: float,
                   nabc: int, free_surface: bool,
                   src_x: torch.Tensor, src_z: torch.Tensor, src_n: int, src_v: torch.Tensor,
                   rcv_x: torch.Tensor, rcv_z: torch.Tensor, rcv_n: int,
                   damp: torch.Tensor,
                   v: torch.Tensor, rho: torch.Tensor,
                   checkpoint_segments: int = 1,
                   device: torch.device = torch.device("cuda"), dtype: torch.dtype = torch.float32
                   ) -> Dict[str, torch.Tensor]:
    """Forward simulation of the acoustic wave equation; same contract as the reference's
    ``forward_kernel`` (acoustic_kernels.py:179-301): returns the dict with keys
    ``p,u,w`` (ns,nt,nr) and ``forward_wavefield_p/u/w`` (nz,nx)."""
    if dtype != torch.float32:
        raise TypeError("adfwi_b200: only torch.float32 is supported (the reference pads to fp32 regardless, "
                        "acoustic_kernels.py:20)")
    _require_cuda(v, rho, damp, src_v)
    if src_v.dim() != 2:
        raise ValueError("adfwi_b200: src_v must be (src_n, nt)")
    if tuple(v.shape) != (nz, nx) or tuple(rho.shape) != (nz, nx):
        raise ValueError(f"adfwi_b200: v/rho must have shape ({nz},{nx})")
    if tuple(damp.shape) != (nz + 2 * nabc, nx + 2 * nabc):
        raise ValueError("adfwi_b200: damp must have the padded shape (nz+2*nabc, nx+2*nabc)")
    if src_v.shape[0] != src_n or src_v.shape[1] != nt:
        raise ValueError("adfwi_b200: src_v must be (src_n, nt)")
    dev = v.device
    src_x, src_z = src_x.to(dev), src_z.to(dev)
    rcv_x, rcv_z = rcv_x.to(dev), rcv_z.to(dev)
    _DeferredChecks.poll()
    _check_indices([("src_x", src_x, nx), ("src_z", src_z, nz), ("rcv_x", rcv_x, nx), ("rcv_z", rcv_z, nz)])
    alpha1, alpha2, kappa1, kappa2, kappa3 = coefficient_planes(v, rho, damp.to(dev).float(), dt, dz, nabc, free_surface)
    rcv_p, rcv_u, rcv_w, ill_p, ill_u, ill_w = AcousticFD.apply(
        alpha1, alpha2, kappa1, kappa2, kappa3, src_v.to(dev),
        src_x + nabc, src_z + nabc, rcv_x + nabc, rcv_z + nabc,
        int(nabc), bool(free_surface), float(dt), int(checkpoint_segments))
    return {
        "p": rcv_p, "u": rcv_u, "w": rcv_w,
        "forward_wavefield_p": ill_p, "forward_wavefield_u": ill_u, "forward_wavefield_w": ill_w,
    }
