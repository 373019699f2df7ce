#!/usr/bin/env python
"""C1: how long does the HOST need to enqueue one gradient (no sync) against the device time of the gradient?  Profiler table of the host side."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from adfwi_b200 import fwi, synthetic as syn
from adfwi_b200.propagator import AcousticPropagator

wl = bench.WORKLOADS["C1"]
dev = torch.device("cuda:0")
nz, nx, nabc, nt, dt, dx = wl["nz"], wl["nx"], wl["nabc"], wl["nt"], wl["dt"], wl["dx"]
ns = 40
vp_true = syn.marmousi_like_vp(nz, nx); vp_init = syn.smooth2d(vp_true, 6)
survey = syn.surface_survey(nx, ns, wl["nr"], nt, dt, wl["f0"])
model = syn.AcousticGridModel(vp_init, dx=dx, dz=dx, nabc=nabc, free_surface=True, vp_grad=True, device=dev)
prop = AcousticPropagator(model, survey, device=dev)
with torch.no_grad():
    obs = AcousticPropagator(syn.AcousticGridModel(vp_true, dx=dx, dz=dx, nabc=nabc, free_surface=True, vp_grad=False, device=dev), survey, device=dev).forward()["p"].clone()
def step():
    model.vp.grad = None
    return fwi.acoustic_gradient(prop, obs, batch_size=ns)
for i in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    step()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"step {i}: host enqueue {1e3 * (t1 - t0):.1f} ms, until device idle {1e3 * (t2 - t0):.1f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=12, max_name_column_width=50))
