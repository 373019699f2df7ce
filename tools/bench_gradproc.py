"""Throughput of the device-side gradient post-processing on the C2 model grid (350x1700) next to the numpy/scipy
oracle (= what the reference runs on the host).  Prints one JSON line; run on the GPU box.
    python tools/bench_gradproc.py > gpurun_out/gradproc_bench.json"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adfwi_b200.propagator import GradProcessor  # noqa: E402
from oracle import gradproc_oracle as GO  # noqa: E402

nz, nx = 350, 1700
rng = np.random.default_rng(0)
grad = rng.standard_normal((nz, nx)).astype(np.float32)
forw = (np.abs(rng.standard_normal((nz, nx))) * np.exp(-np.linspace(0, 5, nz))[:, None] * 50).astype(np.float32)
kw = dict(grad_mute=20, grad_smooth=5, norm_grad=True, forw_illumination=True, marine_or_land="land")
dev = torch.device("cuda:0")
gp = GradProcessor(**kw)
gt, ft = torch.tensor(grad, device=dev), torch.tensor(forw, device=dev)
for _ in range(3):
    out = gp.forward(nx=nx, nz=nz, vmax=np.float32(4700.0), grad=gt, forw=ft)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20
e0.record()
for _ in range(K):
    out = gp.forward(nx=nx, nz=nz, vmax=np.float32(4700.0), grad=gt, forw=ft)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
t0 = time.perf_counter()
out_h = gp.forward(nx=nx, nz=nz, vmax=np.float32(4700.0), grad=grad.copy(), forw=forw)      # host arrays in/out (reference-facing call)
ms_host = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter()
ref = GO.grad_process(nx, nz, np.float32(4700.0), grad.copy(), forw=forw, **kw)
cpu_s = time.perf_counter() - t0
err = float(np.abs(out_h - ref).max() / np.abs(ref).max())
# algorithmic bytes: three smooth2d (taper, illumination, gradient) = 2 passes x (8 B read + 8 B write) each, plus ~10 elementwise passes of 16 B
plane = nz * nx
alg = (3 * 2 * 16 + 10 * 16) * plane
print(json.dumps({"what": "GradProcessor.forward on the C2 model grid 350x1700 (land taper 20, illumination span 40, smoothing span 5, normalise)",
                  "gpu_ms_device_resident": ms, "gpu_ms_host_arrays": ms_host, "cpu_oracle_s": cpu_s, "cpu_cores": 1,
                  "speedup_device_resident": cpu_s * 1e3 / ms, "rel_err_vs_oracle": err,
                  "algorithmic_bytes": alg, "achieved_GBs": alg / (ms * 1e-3) / 1e9,
                  "note": "a 4.8 MB float64 plane: the step is launch-latency bound (about 30 small launches), not HBM bound"}))
