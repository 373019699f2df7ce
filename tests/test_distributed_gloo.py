"""N>1 path on CPU: two gloo ranks shard the shots, each produces the gradient of its shard
(the oracle stands in for the GPU kernels -- this test is about the sharding + all-reduce host
logic), one all-reduce of the flat buffer must reproduce the single-rank gradient."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _shard_gradient(lo, hi):
    from oracle import oracle as O
    g = np.load(os.path.join(ROOT, "tests", "golden", "acoustic_fs.npz"))
    nabc, fs, dt, dz = int(g["nabc"]), bool(g["free_surface"]), float(g["dt"]), float(g["dz"])
    coef = O.acoustic_coefficients(g["vp"], g["rho"], g["damp"], dt, dz, nabc, fs)
    sl = slice(lo, hi)
    out = O.acoustic_run(coef, nabc, fs, dt, g["src_x"][sl], g["src_z"][sl], g["src_v"][sl], g["rcv_x"], g["rcv_z"],
                         g_rcv=(g["W_p"][sl], None, None), illum=True)
    gv, _ = O.acoustic_model_gradients(coef, out["g_alpha1"], out["g_alpha2"], dt, dz, nabc)
    loss = float((out["p"].astype(np.float64) * g["W_p"][sl]).sum())
    return gv.astype(np.float32), out["illum_p"].copy(), loss


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from adfwi_b200 import distributed as D
    r, _, w = D.init_from_env("gloo")
    lo, hi = D.shard_shots(2, r, w)
    gv, illum, loss = _shard_gradient(lo, hi)
    p = torch.nn.Parameter(torch.zeros(gv.shape))
    p.grad = torch.tensor(gv)
    it, lt = torch.tensor(illum), torch.tensor([loss], dtype=torch.float32)
    D.allreduce_gradients([p], extras=[it, lt])
    if r == 0:
        q.put((p.grad.numpy(), it.numpy(), float(lt)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gv2, ill2, loss2 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    gv1, ill1, loss1 = _shard_gradient(0, 2)
    assert np.linalg.norm(gv2 - gv1) / np.linalg.norm(gv1) < 2e-6
    assert np.linalg.norm(ill2 - ill1) / np.linalg.norm(ill1) < 2e-6
    assert abs(loss2 - loss1) <= 1e-5 * abs(loss1)
