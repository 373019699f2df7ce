#!/usr/bin/env python
"""Where does the non-kernel time of a short elastic gradient go?  Prints allocator statistics per step and a torch-profiler table."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from adfwi_b200 import fwi, synthetic as syn
from adfwi_b200.propagator import ElasticPropagator

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C4"]
nt, ns = 400, 15
dev = torch.device("cuda:0")
nz, nx, nabc, dt, dx = wl["nz"], wl["nx"], wl["nabc"], wl["dt"], wl["dx"]
vp_true, vp_init, mk_vs, mk_rho, eps, delta = bench.elastic_fields(wl)
grads = ("eps", "delta") if wl.get("vti") else ("vp", "vs", "rho")
survey = syn.surface_survey(nx, ns, wl["nr"], nt, dt, wl["f0"], src_z=wl["z_sr"], rcv_z=wl["z_sr"])
mk = lambda vp, req: syn.ElasticGridModel(vp, mk_vs(vp), mk_rho(vp), eps=eps, delta=delta, dx=dx, dz=dx, nabc=nabc, free_surface=True, abc_type="PML", requires_grad=req, device=dev)
model = mk(vp_init, grads)
prop = ElasticPropagator(model, survey, device=dev)
with torch.no_grad():
    o = ElasticPropagator(mk(vp_true, ()), survey, device=dev).forward()
    obs = {c: o[c].clone() for c in ("vx", "vz")}
del o
params = [getattr(model, k) for k in grads]
def step():
    for p in params: p.grad = None
    return fwi.elastic_gradient(prop, obs, batch_size=ns)
for i in range(6):
    st0 = torch.cuda.memory_stats()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    step()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    st1 = torch.cuda.memory_stats()
    print(f"step {i}: {1e3 * (t1 - t0):.1f} ms, cudaMalloc calls +{st1['num_device_alloc'] - st0['num_device_alloc']}, cudaFree +{st1['num_device_free'] - st0['num_device_free']}, "
          f"retries +{st1['num_alloc_retries'] - st0['num_alloc_retries']}, reserved {st1['reserved_bytes.all.current'] / 1e9:.1f} GB")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=14, max_name_column_width=60))
