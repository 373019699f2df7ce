"""N>1 path on GPUs: two NCCL ranks (one process per GPU) shard the shots of the reference's acoustic example and of a small
elastic survey, each runs the CUDA propagators on its shard, ONE all-reduce of the flat gradient buffer; rank 0 then repeats the
whole survey alone and the two gradients must agree (they differ only in the order of the fp32 sums over shots).  Needs two
GPUs in one box (`gpurun --gpus 2`); skipped elsewhere."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _acoustic(dev, lo, hi):
    from adfwi_b200 import fwi, synthetic as syn
    from adfwi_b200.propagator import AcousticPropagator
    g = np.load(os.path.join(ROOT, "tests", "golden", "acoustic_c1_scale.npz"))
    nt, dt, f0 = int(g["nt"]), float(g["dt"]), float(g["f0"])
    model = syn.AcousticGridModel(g["vp_init"], rho=g["rho_init"], dx=float(g["dx"]), dz=float(g["dz"]), nabc=int(g["nabc"]),
                                  free_surface=True, vp_grad=True, rho_grad=True, auto_update_rho=False, device=dev)
    prop = AcousticPropagator(model, syn.Survey(syn.Source(np.stack([g["src_x"], g["src_z"]], 1), g["wavelet"], nt, dt, f0),
                                                syn.Receiver(np.stack([g["rcv_x"], g["rcv_z"]], 1))), device=dev)
    prop.damp = torch.tensor(g["damp"], device=dev)
    obs = torch.tensor(g["obs_p"][lo:hi], device=dev)
    loss, illum = fwi.acoustic_gradient(prop, obs, shots=np.arange(lo, hi), batch_size=2)
    return [model.vp, model.rho], [illum.clone(), loss.reshape(1).clone()]


def _elastic(dev, lo, hi):
    from adfwi_b200 import fwi, synthetic as syn
    from adfwi_b200.propagator import ElasticPropagator
    nz, nx, nt, dt, ns = 60, 150, 1000, 1e-3, 4      # long enough for every trace to carry signal (sqrt'(0) of an all-zero residual is NaN upstream too)
    vp = syn.marmousi_like_vp(nz, nx)
    vs, rho = (vp / np.sqrt(3.0)).astype(np.float32), syn.gardner_rho(vp)
    mk = lambda a, b, c: syn.ElasticGridModel(a, b, c, dx=10.0, dz=10.0, nabc=20, free_surface=True, device=dev)
    survey = syn.surface_survey(nx, ns, 50, nt, dt, 15.0, src_z=2, rcv_z=2)
    with torch.no_grad():
        true = ElasticPropagator(mk(vp, vs, rho), survey, device=dev).forward(shot_index=np.arange(lo, hi))
        obs = {c: true[c].clone() for c in ("vx", "vz")}
    model = mk(syn.smooth2d(vp, 4), syn.smooth2d(vs, 4), rho)
    prop = ElasticPropagator(model, survey, device=dev)
    loss, illum = fwi.elastic_gradient(prop, obs, shots=np.arange(lo, hi), batch_size=2)
    return [model.vp, model.vs, model.rho], [illum.clone(), loss.reshape(1).clone()]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from adfwi_b200 import distributed as D
    r, local, w = D.init_from_env("nccl")
    dev = torch.device("cuda", local)
    out = {}
    for name, fn in (("acoustic", _acoustic), ("elastic", _elastic)):
        lo, hi = D.shard_shots(4, r, w)
        params, extras = fn(dev, lo, hi)
        D.allreduce_gradients(params, extras=extras)
        torch.cuda.synchronize(dev)
        if r == 0:
            single_p, single_e = fn(dev, 0, 4)
            out[name] = dict(g=[_rel(a.grad.cpu().numpy(), b.grad.cpu().numpy()) for a, b in zip(params, single_p)],
                             illum=_rel(extras[0].cpu().numpy(), single_e[0].cpu().numpy()),
                             loss=abs(float(extras[1]) - float(single_e[1])) / abs(float(single_e[1])),
                             backend=dist.get_backend())
    if r == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one box")
def test_two_nccl_ranks_equal_single_rank():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    print("2 NCCL ranks vs 1 rank:", out)
    for name, o in out.items():
        assert o["backend"] == "nccl"
        assert max(o["g"]) <= 5e-6, (name, o)
        assert o["illum"] <= 5e-6 and o["loss"] <= 5e-6, (name, o)
