"""Shot-parallel data parallelism: one process per GPU, every rank holds the full model and runs
its own contiguous block of shots; the model gradients (plus illumination and loss) are summed
with ONE all-reduce on a flat fp32 buffer (NCCL over NVLink/NVSwitch on the GPU box, gloo in the
CPU tests).  There is no data-path collective inside the propagation: shots never interact
(batch dim 0 of every array, acoustic_kernels.py:240-242)."""
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_shots(n_shots: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo,hi) of rank ``rank``; the first ``n_shots % world`` ranks get one extra."""
    base, extra = divmod(n_shots, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """torchrun-style initialisation (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    import os
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)      # binds the communicator to this rank's GPU (barrier() then needs no guess)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local, world


def allreduce_sum_(tensors: Iterable[Optional[torch.Tensor]]) -> List[Optional[torch.Tensor]]:
    """In-place SUM over ranks of all given tensors through one flat buffer / one collective."""
    ts = [t for t in tensors if t is not None]
    if not ts or not dist.is_initialized() or dist.get_world_size() == 1:
        return ts
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in ts])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for t in ts:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    return ts


def allreduce_gradients(params: Iterable[torch.nn.Parameter], extras: Iterable[Optional[torch.Tensor]] = ()):
    """Sum ``p.grad`` of every parameter that has one, plus ``extras`` (illumination map, loss)."""
    grads = [p.grad for p in params if p is not None and p.grad is not None]
    return allreduce_sum_(list(grads) + list(extras))
