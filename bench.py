#!/usr/bin/env python
"""bench.py -- FWI-gradient throughput of the wave-propagation hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C2]
    torchrun ... bench.py --gpus N ...          (one rank per GPU, NCCL; launched by the driver)

A "step" = one FWI gradient over this rank's shots: for every shot batch forward modelling,
L2 waveform misfit against observed records, adjoint sweep, gradient accumulation; then (N>1) one
all-reduce of the model gradient.  Workload (default C2, BASELINE.json configs[1]): iso-acoustic
350x1700 grid (10 m, 50-cell PML + free surface -> 450x1800 padded), nt=4000, 1700 receivers,
240 shots over 8 GPUs = 30 shots per GPU (weak scaling: per-GPU work fixed).  Synthetic model,
random-free analytic velocity; observed data = our own forward modelling of the
"true" model, generated before the timed region.

metric `value`: 2*nzp*nxp*nt*shots / time  (forward + adjoint cell-updates per second, summed over
GPUs), inputs resident in HBM.  `e2e`: same through AcousticPropagator.forward()+backward() with
the model, wavelets and observed data starting in pinned HOST memory and gradient+loss read back.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nz, nx, dx, dt, nt, f0, nabc, shots_total_at_8gpus, nr, batch_size)
    "C1": dict(nz=88, nx=200, dx=40.0, dt=3e-3, nt=1600, f0=5.0, nabc=30, shots8=40 * 8, nr=200, batch=40,
               desc="iso-acoustic Marmousi2 example 88x200 (148x260 padded), nt=1600"),
    "C2": dict(nz=350, nx=1700, dx=10.0, dt=1e-3, nt=4000, f0=10.0, nabc=50, shots8=240, nr=1700, batch=10,
               desc="iso-acoustic Marmousi2 full-res 350x1700 (450x1800 padded), nt=4000, 240 shots / 8 GPUs"),
    "C3": dict(kind="elastic", nz=350, nx=1700, dx=10.0, dt=1e-3, nt=4000, f0=10.0, nabc=50, shots8=240, nr=1700, batch=30, z_sr=10,
               desc="iso-elastic Marmousi2 350x1700 (402x1800 padded), split-PML O(2,4), free surface, nt=4000, vp/vs/rho gradients, 240 shots / 8 GPUs"),
    "C4": dict(kind="elastic", nz=320, nx=720, dx=2.5, dt=2.5e-4, nt=4000, f0=30.0, nabc=50, shots8=120, nr=720, batch=15, z_sr=10, vti=True,
               desc="VTI-elastic 320x720 (372x820 padded), split-PML O(2,4), free surface, nt=4000, eps/delta gradients, 120 shots / 8 GPUs"),
    # the full-length C5 job (nt=8000: 570 GB of stencil history per shot) can only run checkpointed, so its slice runs
    # that way too: 8 shots per launch, history kept for 125 steps at a time, one recomputation sweep (counted as overhead)
    "C5": dict(nz=2048, nx=8192, dx=5.0, dt=5e-4, nt=1000, f0=15.0, nabc=50, shots8=8 * 8, nr=8192, batch=8, ckpt=125,
               desc="synthetic acoustic 2048x8192 (2148x8292 padded), 1000-step slice of nt=8000, 8 shots/GPU, "
                    "checkpointed every 125 steps as the full-length run must be"),
}
# algorithmic bytes per cell-update
B_FWD_SAVE = 36.0     # forward sweep in recording mode: p,u,w r+w 24 + alpha1,alpha2 8 + S write 4
B_ADJ = 44.0          # adjoint sweep (vp only): 3 adjoint fields r+w 24 + alpha1,alpha2 8 + S read 4 + g_alpha1 RMW 8
B_GRAD_STEP = 80.0    # forward(recording) + adjoint per cell-step = 40 B per cell-update
B_FWD_PLAIN = 32.0    # forward sweep without recording (the recomputation-free part of a checkpointed run)
# with the density gradient ("with rho"): two more history planes written / read, g_alpha2 read-modify-write
B_FWD_SAVE_RHO = 44.0
B_ADJ_RHO = 60.0
L2_BYTES = 126e6


def l2_note(state_bytes_per_launch, stream_bytes_per_step):
    """config.l2: what the timed region does about the L2 (timing rule: flush it or use inputs larger than it, and say which).
    `state_bytes_per_launch` = the wavefield planes one launch WRITES and the next launch (next time step, same shots) reads."""
    fits = state_bytes_per_launch <= L2_BYTES
    return {"flush": "none (no explicit flush)",
            "bytes_streamed_per_step": int(stream_bytes_per_step),
            "inputs_larger_than_l2": bool(stream_bytes_per_step > 4 * L2_BYTES),
            "state_bytes_handed_between_launches": int(state_bytes_per_launch),
            "state_fits_l2": bool(fits),
            "note": ("every timed step streams its stencil history, records and receiver cotangents through HBM (far more than the 126 MB L2), so "
                     "nothing of one step survives in L2 into the next; " +
                     ("WITHIN a step the wavefield state of one launch fits in L2 and is re-read from there by the next time step's launch -- "
                      "a property of the algorithm the kernels exploit, which is why measured DRAM traffic is below the algorithmic bytes"
                      if fits else "the wavefield state of one launch is larger than L2 too: consecutive launches stream it from HBM"))}
# elastic split-PML: forward 104 B (+20 B recording), adjoint 124 B (+ gradient RMW amortised over the shots)
B_EL_FWD_SAVE = 124.0
B_EL_ADJ = 124.0
B_EL_GRAD_STEP = 248.0


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        if os.environ.get("ADFWI_BENCH_NO_CLOCKS"):        # diagnostics only: does the sampler itself perturb a short timed region?
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); smax = max(smax, float(f[1]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": smax, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle (port of the reference time loop + its adjoint)
# ------------------------------------------------------------------------------------------------
def cpu_gradient_sample(wl, ns, nt):
    """Time forward+adjoint of the oracle on a bounded sample (ns shots x nt steps) of the
    workload's grid with all host threads.  Returns (cell_updates_per_s, seconds)."""
    from oracle import oracle as O
    from adfwi_b200 import synthetic as syn
    from adfwi_b200.propagator.boundary_condition import bc_pml
    if wl.get("kind") == "elastic":
        return cpu_elastic_sample(wl, ns, nt)
    nz, nx, nabc = wl["nz"], wl["nx"], wl["nabc"]
    vp = syn.smooth2d(syn.marmousi_like_vp(nz, nx), 6)
    rho = syn.gardner_rho(vp)
    damp = bc_pml(nx, nz, wl["dx"], wl["dx"], pml=nabc, vmax=float(vp.max()), free_surface=False).astype(np.float32)
    coef = O.acoustic_coefficients(vp, rho, damp, wl["dt"], wl["dx"], nabc, True)
    sx = np.round(np.linspace(2, nx - 3, ns)).astype(np.int64); sz = np.ones(ns, np.int64)
    rx = np.round(np.linspace(0, nx - 1, wl["nr"])).astype(np.int64); rz = np.ones(wl["nr"], np.int64)
    wav = np.broadcast_to(syn.integrated_ricker(nt, wl["dt"], wl["f0"] * 4).astype(np.float32), (ns, nt)).copy()
    O.lib().oracle_set_threads(host_cores())     # torchrun exports OMP_NUM_THREADS=1
    t0 = time.perf_counter()
    # one forward sweep (with history), residual-like cotangent from the records, one adjoint sweep
    O.acoustic_run(coef, nabc, True, wl["dt"], sx, sz, wav, rx, rz, g_rcv=lambda rec: (rec["p"], None, None),
                   need_g_alpha2=False)
    sec = time.perf_counter() - t0
    cells = (nz + 2 * nabc) * (nx + 2 * nabc) * ns * nt
    return 2.0 * cells / sec, sec


def elastic_fields(wl):
    """Synthetic iso / VTI elastic model of a workload: (true vp, initial vp, vs, rho, eps, delta) as numpy planes."""
    from adfwi_b200 import synthetic as syn
    nz, nx = wl["nz"], wl["nx"]
    vp_true = syn.marmousi_like_vp(nz, nx)
    vp = syn.smooth2d(vp_true, 6)
    mk = lambda v: (v / np.sqrt(3.0)).astype(np.float32)
    eps = np.full((nz, nx), 0.1 if wl.get("vti") else 0.0, np.float32)
    delta = np.full((nz, nx), -0.1 if wl.get("vti") else 0.0, np.float32)
    return vp_true, vp, mk, syn.gardner_rho, eps, delta


def cpu_elastic_sample(wl, ns, nt):
    """Elastic counterpart of cpu_gradient_sample: the oracle's split-PML forward + adjoint."""
    from oracle import oracle as O
    from adfwi_b200 import synthetic as syn
    from adfwi_b200.propagator.boundary_condition import bc_pml_xz
    nz, nx, nabc, dx = wl["nz"], wl["nx"], wl["nabc"], wl["dx"]
    _, vp, mk_vs, mk_rho, eps, delta = elastic_fields(wl)
    vs, rho = mk_vs(vp), mk_rho(vp)
    C33 = vp * vp * rho; C44 = vs * vs * rho; C11 = C33 * (1 + 2 * eps)
    C13 = np.sqrt(2 * C33 * (C33 - C44) * delta + (C33 - C44) ** 2) - C44
    b = 1.0 / rho
    C55 = 0.2 * (C44[1:-1, 1:-1] + C44[2:, 1:-1] + C44[1:-1, 2:] + C44[2:, 1:-1] + C44[2:, 2:])
    planes = dict(C11=C11, C13=C13, C33=C33, C55=C55, bx=0.5 * (b[:, :-1] + b[:, 1:]), bz=0.5 * (b[:-1] + b[1:]))
    planes = {k: v.astype(np.float32) for k, v in planes.items()}
    bcx, bcz = bc_pml_xz(nx, nz, dx, dx, pml=nabc, vmax=float(vp.max()), free_surface=True)
    z = wl["z_sr"]
    sx = np.round(np.linspace(2, nx - 3, ns)).astype(np.int64); sz = np.full(ns, z, np.int64)
    rx = np.round(np.linspace(0, nx - 1, wl["nr"])).astype(np.int64); rz = np.full(wl["nr"], z, np.int64)
    wav = np.broadcast_to(syn.integrated_ricker(nt, wl["dt"], wl["f0"] * 4).astype(np.float32), (ns, nt)).copy()
    mt = np.broadcast_to(np.eye(3, dtype=np.float32), (ns, 3, 3)).copy()
    O.lib().oracle_set_threads(host_cores())
    t0 = time.perf_counter()
    O.elastic_run(planes, "PML", 4, True, nz, nx, nabc, dx, dx, wl["dt"], sx, sz, wav, mt, rx, rz, bcx=bcx.astype(np.float32),
                  bcz=bcz.astype(np.float32), g_rcv=lambda rec: (None, None, None, rec["vx"], rec["vz"]))
    sec = time.perf_counter() - t0
    cells = (nz + nabc + 2) * (nx + 2 * nabc) * ns * nt
    return 2.0 * cells / sec, sec


def host_cores():
    """Host threads the CPU legs may use: the smallest of os.cpu_count(), the scheduler affinity mask and the cgroup CPU quota.  A GPU
    box that is a slice of a larger node can SHOW more CPUs than its quota lets run at once; an OpenMP / ATen thread team larger than
    the quota stalls at every parallel-region barrier (seen: the unmodified reference 80x slower with 24 threads on such a slice)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:                       # cgroup v2: "<quota|max> <period>"
            q, per = f.read().split()[:2]
        if q != "max":
            n = min(n, max(1, int(int(q) / int(per))))
    except (OSError, ValueError):
        try:
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f:      # cgroup v1
                q = int(f.read())
            with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
                per = int(f.read())
            if q > 0:
                n = min(n, max(1, q // per))
        except (OSError, ValueError):
            pass
    return max(1, n)


def cpu_sample_size():
    """Bounded CPU sample: the oracle's adjoint parallelises over shots, so give it one shot per
    host thread (up to 16) for 50 time steps."""
    return max(2, min(host_cores(), 16)), 50


def cpu_port_warm(wl, reps=2):
    """Warm mean of the oracle port on the bounded sample (first call discarded: page faults of the history, OpenMP start-up)."""
    ns, nt = cpu_sample_size()
    cpu_gradient_sample(wl, ns, nt)
    vals, secs = zip(*[cpu_gradient_sample(wl, ns, nt) for _ in range(reps)])
    return float(np.mean(vals)), float(np.mean(secs)), ns, nt


def reference_objects(wl, ns, nt, device):
    """The UNMODIFIED reference's own model / survey / propagator objects for an (ns shots x nt steps) sample of a
    workload, on `device` ('cpu' or 'cuda:0').  Needs the reference package (oracle/ref_loader.py search path)."""
    from oracle import ref_loader
    from adfwi_b200 import synthetic as syn
    ref_loader.load()
    from ADFWI.survey import Source, Receiver, Survey
    nz, nx, nabc, dx, dt, f0 = wl["nz"], wl["nx"], wl["nabc"], wl["dx"], wl["dt"], wl["f0"]
    z = wl.get("z_sr", 1)
    sx = np.round(np.linspace(2, nx - 3, ns)).astype(int) if ns > 1 else np.array([nx // 2])
    rx = np.round(np.linspace(0, nx - 1, wl["nr"])).astype(int)
    s = Source(nt=nt, dt=dt, f0=f0)
    s.add_sources(src_x=sx, src_z=np.full(ns, z), src_wavelet=syn.integrated_ricker(nt, dt, f0 * 4).astype(np.float32), src_type="mt", src_mt=np.eye(3))
    r = Receiver(nt=nt, dt=dt)
    r.add_receivers(rcv_x=rx, rcv_z=np.full(len(rx), z), rcv_type="pr")
    survey = Survey(source=s, receiver=r)
    vp = syn.smooth2d(syn.marmousi_like_vp(nz, nx), 6)
    if wl.get("kind") == "elastic":
        from ADFWI.model import AnisotropicElasticModel, IsotropicElasticModel
        from ADFWI.propagator import ElasticPropagator
        vs, rho = (vp / np.sqrt(3.0)).astype(np.float32), syn.gardner_rho(vp)
        if wl.get("vti"):
            one = np.ones((nz, nx), np.float32)
            model = AnisotropicElasticModel(0, 0, nx, nz, dx, dx, vp=vp, vs=vs, rho=rho, eps=0.1 * one, gamma=0 * one, delta=-0.1 * one,
                                            eps_grad=True, delta_grad=True, free_surface=True, anisotropic_type="vti", abc_type="PML",
                                            nabc=nabc, device=device)
        else:
            model = IsotropicElasticModel(0, 0, nx, nz, dx, dx, vp, vs, rho, vp_grad=True, vs_grad=True, rho_grad=True, free_surface=True,
                                          abc_type="PML", nabc=nabc, auto_update_rho=False, auto_update_vp=False, device=device)
        return model, ElasticPropagator(model, survey, device=device), ("vx", "vz")
    from ADFWI.model import AcousticModel
    from ADFWI.propagator import AcousticPropagator
    model = AcousticModel(0, 0, nx, nz, dx, dx, vp, syn.gardner_rho(vp), vp_grad=True, free_surface=True, abc_type="PML", nabc=nabc, device=device)
    return model, AcousticPropagator(model, survey, device=device), ("p",)


def reference_gradient_sample(wl, ns, nt, device, reps, warmup, threads=None):
    """Time forward + L2 misfit + loss.backward() of the unmodified reference (its own autograd tape through the TorchScript time
    loop, checkpoint_segments=4 as in its examples) on an ns x nt sample.  Returns (cell-updates/s, seconds per gradient)."""
    import torch
    from oracle import ref_loader
    ref_loader.load()
    from ADFWI.fwi.misfit import Misfit_waveform_L2
    torch.set_num_threads(threads or host_cores())    # torchrun exports OMP_NUM_THREADS=1
    model, prop, comps = reference_objects(wl, ns, nt, device)
    fn = Misfit_waveform_L2(dt=wl["dt"])
    elastic = wl.get("kind") == "elastic"
    with torch.no_grad():
        o = prop.forward()
        obs = {c: 0.9 * o[c].detach() for c in comps}
    secs = []
    for i in range(warmup + reps):
        for p in model.parameters():
            p.grad = None
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        rec = prop.forward(fd_order=4, checkpoint_segments=4) if elastic else prop.forward(checkpoint_segments=4)
        loss = sum(fn.forward(obs[c], rec[c]) for c in comps)
        loss.backward()
        if device != "cpu":
            torch.cuda.synchronize()
        if i >= warmup:
            secs.append(time.perf_counter() - t0)
    nzp = wl["nz"] + (wl["nabc"] + 2 if elastic else 2 * wl["nabc"])
    cells = nzp * (wl["nx"] + 2 * wl["nabc"]) * ns * nt
    sec = float(np.mean(secs))
    return 2.0 * cells / sec, sec


def cpu_baseline_entry(args, wl):
    """`cpu_baseline` of the B200 arm: what the reference arm measures -- the unmodified reference on the host cores when its package is
    staged (kind "reference"), else the C/OpenMP port (kind "port"), the port always beside it -- taken in a FRESH process (this one
    carries a CUDA context, torch's thread pools and the clock sampler, which halved the port's throughput when it ran in here)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", "1", "--warmup", "1", "--no-reference-cuda"]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
        d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")][-1])
        out = dict(d["cpu_baseline"])
        out["port"] = d.get("port")
        return out
    except Exception as e:
        return {"value": None, "unit": "Gcell-updates/s", "cores": host_cores(), "kind": "port", "sample": f"cpu baseline run failed: {type(e).__name__}: {e}"[:300]}


def run_reference(args, wl):
    """Reference arm: the UNMODIFIED reference (pure PyTorch: TorchScript time loop + autograd tape) on the host cores when its
    package is present (baseline/_ref staged by baseline/stage_reference.py, or /root/reference), timed on a bounded sample of the
    workload; otherwise the C/OpenMP port under oracle/.  The port's number and -- when a GPU is visible -- the reference's own
    CUDA path (device='cuda', the mode all its examples run in) are reported beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_loader
    cores = host_cores()
    port_v, port_s, pns, pnt = cpu_port_warm(wl, reps=max(args.steps, 1))
    port = {"value": port_v / 1e9, "unit": "Gcell-updates/s", "cores": cores, "kind": "port",
            "sample": f"oracle/ C port of the reference time loop + adjoint, OpenMP over {cores} host threads, {pns} shots x {pnt} steps of the "
                      f"{args.workload} grid, warm mean of {max(args.steps, 1)}, {port_s:.1f} s each"}
    elastic = wl.get("kind") == "elastic"
    ref_cuda = None
    cores_used = cores
    if ref_loader.available():
        ns, nt = (2, 20) if elastic else (2, 100)
        try:
            # the reference issues ~2 k small ATen ops per step, each a fork-join over the whole thread team: on a many-core box the
            # widest team is not the fastest one.  Give it the team size it runs best with (one warm gradient each), then time that.
            tried = {}
            # (ATen splits these tensors into at most ~50 chunks, so teams beyond 64 threads only add fork-join cost: not tried)
            for th in sorted({min(cores, 16), min(cores, 32), min(cores, 64)}):      # ascending; stop widening once it no longer pays
                tried[th] = reference_gradient_sample(wl, ns, nt, "cpu", reps=1, warmup=1, threads=th)[1]
                if tried[th] > 1.1 * min(tried.values()):
                    break
            best = min(tried, key=tried.get)
            # keep the whole arm within a few minutes whatever the box: shorten the sample (fewer time steps), never the step count
            total = tried[best] * (max(args.steps, 1) + max(args.warmup, 1))
            if total > 120.0:
                nt = max(10, int(nt * 120.0 / total))
            val, sec = reference_gradient_sample(wl, ns, nt, "cpu", reps=max(args.steps, 1), warmup=max(args.warmup, 1), threads=best)
            kind = "reference"
            cores_used = best
            sample = (f"unmodified reference (ADFWI propagator.forward + Misfit_waveform_L2 + loss.backward(), checkpoint_segments=4, torch "
                      f"{best} threads of {cores} host cores" + (f" [seconds per gradient by team size: {({k: round(v, 2) for k, v in tried.items()})}]" if len(tried) > 1 else "") +
                      f", device='cpu'), {ns} shots x {nt} steps of the {args.workload} grid, {sec:.1f} s per gradient")
        except Exception as e:       # an unusable staging must not take the arm down: fall back to the port
            val, sec, kind, sample = port_v, port_s, "port", port["sample"] + f" (reference failed: {type(e).__name__}: {e})"
        try:
            import torch
            if args.reference_cuda and torch.cuda.is_available():
                cns, cnt = (2, 50) if elastic else (4, 100)
                cv, cs = reference_gradient_sample(wl, cns, cnt, "cuda:0", reps=2, warmup=2)
                ref_cuda = {"value": cv / 1e9, "unit": "Gcell-updates/s", "seconds_per_gradient": cs,
                            "sample": f"unmodified reference with device='cuda:0' (eager ATen kernels + autograd tape), {cns} shots x {cnt} steps of the {args.workload} grid"}
        except Exception as e:
            ref_cuda = {"unavailable": f"{type(e).__name__}: {e}"}
    else:
        val, sec, kind, sample = port_v, port_s, "port", port["sample"]
        ns, nt = pns, pnt
    val_g = val / 1e9
    line = {
        "impl": "reference", "metric": "forward+adjoint cell-updates/s (FWI gradient)", "value": val_g,
        "unit": "Gcell-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "sample": f"{ns} shots x {nt} steps of the same padded grid"},
        "cpu_baseline": {"value": val_g, "unit": "Gcell-updates/s", "cores": cores_used, "kind": kind, "sample": sample},
        "port": port, "reference_cuda": ref_cuda,
        "e2e": {"value": val_g, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args, wl):
    import torch
    import torch.distributed as dist
    from adfwi_b200 import _lib, distributed as D, fwi, synthetic as syn
    from adfwi_b200.propagator import AcousticPropagator
    from adfwi_b200.propagator.acoustic_kernels import config as ak_cfg
    if wl.get("ckpt"):
        ak_cfg["ckpt_interval"] = int(wl["ckpt"])

    rank, local, world = D.init_from_env("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    lib = _lib.load()

    nz, nx, nabc, nt, dt, dx = wl["nz"], wl["nx"], wl["nabc"], wl["nt"], wl["dt"], wl["dx"]
    if args.nt:
        nt = args.nt
    ns_local = args.shots or max(wl["shots8"] // 8, 1)
    ns_total = ns_local * world
    lo, hi = D.shard_shots(ns_total, rank, world)
    batch = min(args.batch or wl["batch"], ns_local)
    nzp, nxp = nz + 2 * nabc, nx + 2 * nabc

    vp_true = syn.marmousi_like_vp(nz, nx)
    vp_init = syn.smooth2d(vp_true, 6)
    survey = syn.surface_survey(nx, ns_total, wl["nr"], nt, dt, wl["f0"])
    true_model = syn.AcousticGridModel(vp_true, dx=dx, dz=dx, nabc=nabc, free_surface=True, vp_grad=False, device=dev)
    model = syn.AcousticGridModel(vp_init, dx=dx, dz=dx, nabc=nabc, free_surface=True, vp_grad=True, rho_grad=args.rho_grad,
                                  auto_update_rho=not args.rho_grad, device=dev)
    b_fwd, b_adj = (B_FWD_SAVE_RHO, B_ADJ_RHO) if args.rho_grad else (B_FWD_SAVE, B_ADJ)
    params = [model.vp, model.rho] if args.rho_grad else [model.vp]
    prop_true = AcousticPropagator(true_model, survey, device=dev)
    prop = AcousticPropagator(model, survey, device=dev)
    prop.damp = prop_true.damp            # same absorbing profile for both (vmax of the true model)
    shots = np.arange(lo, hi)

    # observed data of this rank's shots (untimed set-up), device copy + pinned host copy
    obs = torch.empty((len(shots), nt, wl["nr"]), device=dev)
    with torch.no_grad():
        for pos in fwi.shot_batches(len(shots), batch):
            obs[pos] = prop_true.forward(shot_index=shots[pos])["p"]
    obs_host = obs.cpu().pin_memory()
    vp_host = model.vp.detach().cpu().pin_memory()
    wav_host = prop.wavelet.detach().cpu().pin_memory()
    grad_host = torch.empty((nz, nx), dtype=torch.float32).pin_memory()
    del prop_true, true_model
    torch.cuda.empty_cache()

    def step_resident():
        for p in params:
            p.grad = None
        loss, illum = fwi.acoustic_gradient(prop, obs, shots=shots, batch_size=batch)
        D.allreduce_gradients(params, extras=[illum, loss])
        return loss

    def step_e2e():
        # inputs start in pinned host memory: model, wavelets, observed records of every batch
        for p in params:
            p.grad = None
        with torch.no_grad():
            model.vp.copy_(vp_host, non_blocking=True)
            prop.wavelet.copy_(wav_host, non_blocking=True)
        loader = lambda pos: obs_host[pos[0]:pos[-1] + 1].to(dev, non_blocking=True)
        loss, illum = fwi.acoustic_gradient(prop, None, shots=shots, batch_size=batch, obs_loader=loader)
        D.allreduce_gradients(params, extras=[illum, loss])
        grad_host.copy_(model.vp.grad, non_blocking=True)
        return float(loss.item())      # device->host read of the loss (synchronises)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        trace = [time.perf_counter()]
        st0 = torch.cuda.memory_stats(dev) if os.environ.get("ADFWI_BENCH_TRACE") else None
        for _ in range(steps):
            fn()
            trace.append(time.perf_counter())
        e1.record()
        barrier()
        if st0 is not None:      # diagnostics: host-side time at which each step's calls returned; allocator traffic to the driver
            st1 = torch.cuda.memory_stats(dev)
            keys = ("num_device_alloc", "num_device_free", "num_alloc_retries", "num_ooms")
            sys.stderr.write(f"[trace] {fn.__name__}: " + " ".join(f"{(b - a) * 1e3:.1f}" for a, b in zip(trace, trace[1:])) + " ms; " +
                             ", ".join(f"{k} +{st1.get(k, 0) - st0.get(k, 0)}" for k in keys) + "\n")
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_resident()
    barrier()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start(); time.sleep(0.3)
    n0 = _lib.launch_count()
    _lib.timing_collect()
    _lib.timing_enable(max(16 * args.steps, 16))
    t_wall0 = time.time()
    ms = timed(step_resident, args.steps)
    t_wall1 = time.time()
    _lib.timing_enable(0)
    launches = _lib.launch_count() - n0
    kt = _lib.timing_collect()
    clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None

    for _ in range(min(args.warmup, 3)):      # the first end-to-end steps of a process run slow now and then (pinned staging, allocator growth)
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    updates_per_step = 2.0 * nzp * nxp * nt * ns_total
    value = updates_per_step * args.steps / (ms * 1e-3) / 1e9
    e2e_value = updates_per_step * args.steps / (ms_e2e * 1e-3) / 1e9
    lt = torch.tensor([float(launches)], device=dev)
    if world > 1:
        dist.all_reduce(lt)
    h2d = vp_host.numel() * 4 + wav_host.numel() * 4 + obs_host.numel() * 4
    d2h = grad_host.numel() * 4 + 4

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        # dominant kernel (pair): sampled average launch durations from CUDA events on the launch stream
        avg = {k: v[0] / v[1] for k, v in kt.items()}
        G_cells = None
        roof = None
        if avg:
            adj = sum(avg.get(k, 0.0) for k in ("ac_adj_a", "ac_adj_b", "ac_adj_inject", "ac_adj_fused"))
            fwd = sum(avg.get(k, 0.0) for k in ("ac_fwd_p", "ac_fwd_uw", "ac_record", "ac_fwd_fused"))
            persist = "ac_fwd_persist" in avg        # small grid: the whole sweep of a batch is ONE launch (acp_fwd / acp_adj)
            if persist:
                fwd, adj = avg["ac_fwd_persist"], avg.get("ac_adj_persist", 0.0)
            # cells one launch processes: shots of one library shot group x padded plane
            import ctypes
            from adfwi_b200.propagator.acoustic_kernels import config as ak_config, make_desc
            d = make_desc(nzp, nxp, batch, nt, wl["nr"], nabc, True, dt, 1, True, 0, args.rho_grad, ak_config["shots_per_group"])
            G = lib.adfwi_acoustic_group_size(ctypes.byref(d))   # shots one launch advances
            G_cells = G * nzp * nxp
            if persist:
                G, G_cells = batch, batch * nzp * nxp * nt
            fused = "ac_adj_fused" in avg or "ac_fwd_fused" in avg
            dom_name, dom_ms, dom_bytes, dom_key = \
                (("adjoint step (ac_adj_fused)" if fused else "adjoint step (ac_adj_inject+ac_adj_a+ac_adj_b)"), adj, b_adj, "ac_adj_fused") if adj >= fwd else \
                (("forward step, recording (ac_fwd_fused)" if fused else "forward step (ac_fwd_p+ac_fwd_uw+ac_record)"), fwd, b_fwd, "ac_fwd_fused")
            # measured DRAM bytes per launch of that kernel from the committed `ncu --set full` capture of this command
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            if os.path.exists(tpath):
                try:
                    tj = json.load(open(tpath))
                    wkey = args.workload + ("_rho" if args.rho_grad else "")
                    ent = tj.get(wkey) if isinstance(tj.get(wkey), dict) else (tj if (tj.get("workload") == wkey) else None)
                    if persist:
                        dom_key = "acp_adj" if adj >= fwd else "acp_fwd"
                    if ent and ent.get("batch") == batch:
                        traffic = ent.get(dom_key)
                except Exception:
                    traffic = None
            if persist:
                dom_name = ("adjoint sweep, all time steps in one launch (acp_adj)" if adj >= fwd else "forward sweep, recording, all time steps in one launch (acp_fwd)")
            if dom_ms > 0:
                ach = dom_bytes * G_cells / (dom_ms * 1e-3) / 1e9
                roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "traffic": traffic, "kernel": dom_name, "avg_launch_ms": dom_ms,
                        "algorithmic_bytes_per_cell_update": dom_bytes, "cells_per_launch": G_cells,
                        "peak_source": f"{peak_src} (MEASURED_PEAKS.json hbm_gbs)",
                        # whole gradient against the store-all row (forward recording + adjoint) and, when the history does not fit
                        # and a recomputation sweep runs (ckpt_interval set), against the checkpointed row (+ one plain forward sweep)
                        "whole_step_frac": (b_fwd + b_adj) / 2 * value * 1e9 / world / (peak * 1e9),
                        "whole_step_frac_checkpointed_row": ((b_fwd + b_adj + B_FWD_PLAIN) / 2 * value * 1e9 / world / (peak * 1e9)) if wl.get("ckpt") else None,
                        "kernel_share_of_step": (fwd + adj) * (1 if persist else nt) * (ns_local / max(G, 1)) / (ms / args.steps) if (fwd > 0 and adj > 0 and not wl.get("ckpt")) else None,
                        "persistent": ("cluster-persistent small-grid kernels: the wavefield state stays in shared memory for the whole time loop, HBM sees the "
                                       "stencil history and the records only (8 B per cell-update), so the fraction of the per-step algorithmic-byte roofline "
                                       "can exceed 1") if persist else None,
                        "frac_by_sweep": {"forward_recording": b_fwd * G_cells / (fwd * 1e-3) / 1e9 / peak if fwd > 0 else None,
                                          "adjoint": b_adj * G_cells / (adj * 1e-3) / 1e9 / peak if adj > 0 else None},
                        "note": "algorithmic bytes are SURVEY.md 8(d)'s per-cell-update figures, which count the coefficient planes and "
                                "the gradient read-modify-write once per shot; the fused kernels keep both on chip for the shots of a "
                                "tile walk and the working set of small grids stays in L2, so `achieved` can exceed the DRAM traffic "
                                "actually moved (see `traffic`) and, on long tile walks, the copy-bandwidth `peak`",
                        "per_kernel_avg_ms": avg}
        # N = 1 only: at N > 1 the other ranks wait (spinning host threads) while rank 0 would time a thread team as wide as the box --
        # one busy extra thread stalls every OpenMP barrier of the reference (seen: 40-80x slower at N = 2)
        cpu_base = cpu_baseline_entry(args, wl) if (args.cpu_baseline and world == 1) else None
        line = {
            "metric": "forward+adjoint cell-updates/s (FWI gradient)", "value": value, "unit": "Gcell-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "shots_per_gpu": ns_local, "shots_total": ns_total,
                       "batch_size": batch, "padded_grid": [nzp, nxp], "nt": nt, "receivers": wl["nr"],
                       "gradients": ["vp", "rho"] if args.rho_grad else ["vp"],
                       "l2": l2_note(3.0 * (G * nzp * nxp if avg else batch * nzp * nxp) * 4, (1 + (2 if args.rho_grad else 0)) * 2.0 * nzp * nxp * nt * ns_local * 4),
                       "parallelism": f"shots sharded over {world} GPU(s), one all-reduce of the gradient"},
            "shots_per_s": ns_total * args.steps / (ms * 1e-3),
            "e2e": {"value": e2e_value, "unit": "Gcell-updates/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(lt.item()),
            "roofline": roof,
            "cpu_baseline": cpu_base,
            "clocks": clk,
        }
        if args.secondary and world == 1:
            # the secondary runs are separate processes on the same GPU: hand the memory back first (a child that finds the device
            # full of this process's cached workspace would be pushed into checkpointing and measure a recomputation sweep)
            import gc
            model.vp.grad = None
            del obs, obs_host, prop, model
            gc.collect()
            torch.cuda.empty_cache()
            line["secondary"] = secondary_measurements(args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_b200_elastic(args, wl):
    """C3 / C4: split-PML elastic FWI gradient (vx, vz misfit) through ElasticPropagator.forward() + backward()."""
    import torch
    import torch.distributed as dist
    from adfwi_b200 import _lib, distributed as D, fwi, synthetic as syn
    from adfwi_b200.propagator import ElasticPropagator

    rank, local, world = D.init_from_env("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    _lib.load()

    nz, nx, nabc, nt, dt, dx = wl["nz"], wl["nx"], wl["nabc"], wl["nt"], wl["dt"], wl["dx"]
    if args.nt:
        nt = args.nt
    ns_local = args.shots or max(wl["shots8"] // 8, 1)
    ns_total = ns_local * world
    lo, hi = D.shard_shots(ns_total, rank, world)
    batch = min(args.batch or wl["batch"], ns_local)
    NN = args.order // 2
    nzp, nxp = nz + nabc + NN, nx + 2 * nabc
    comps = ("vx", "vz")
    vp_true, vp_init, mk_vs, mk_rho, eps, delta = elastic_fields(wl)
    grads = ("eps", "delta") if wl.get("vti") else ("vp", "vs", "rho")
    survey = syn.surface_survey(nx, ns_total, wl["nr"], nt, dt, wl["f0"], src_z=wl["z_sr"], rcv_z=wl["z_sr"])
    mk = lambda vp, req: syn.ElasticGridModel(vp, mk_vs(vp), mk_rho(vp), eps=eps, delta=delta, dx=dx, dz=dx, nabc=nabc,
                                              free_surface=True, abc_type=args.abc, requires_grad=req, device=dev)
    true_model, model = mk(vp_true, ()), mk(vp_init, grads)
    prop_true = ElasticPropagator(true_model, survey, device=dev)
    prop = ElasticPropagator(model, survey, device=dev)
    prop.bcx, prop.bcz, prop.damp = prop_true.bcx, prop_true.bcz, prop_true.damp
    shots = np.arange(lo, hi)
    params = [getattr(model, k) for k in grads]

    obs = {c: torch.empty((len(shots), nt, wl["nr"]), device=dev) for c in comps}
    with torch.no_grad():
        for pos in fwi.shot_batches(len(shots), batch):
            rec = prop_true.forward(shot_index=shots[pos], fd_order=args.order)
            for c in comps:
                obs[c][pos] = rec[c]
            del rec
    obs_host = {c: obs[c].cpu().pin_memory() for c in comps}
    par_host = [p.detach().cpu().pin_memory() for p in params]
    wav_host = prop.wavelet.detach().cpu().pin_memory()
    grad_host = [torch.empty((nz, nx), dtype=torch.float32).pin_memory() for _ in params]
    del prop_true, true_model
    torch.cuda.empty_cache()

    def step_resident():
        for p in params:
            p.grad = None
        loss, illum = fwi.elastic_gradient(prop, obs, shots=shots, batch_size=batch, components=comps, fd_order=args.order)
        D.allreduce_gradients(params, extras=[illum, loss])
        return loss

    def step_e2e():
        for p in params:
            p.grad = None
        with torch.no_grad():
            for p, h in zip(params, par_host):
                p.copy_(h, non_blocking=True)
            prop.wavelet.copy_(wav_host, non_blocking=True)
        loader = lambda pos: {c: obs_host[c][pos[0]:pos[-1] + 1].to(dev, non_blocking=True) for c in comps}
        loss, illum = fwi.elastic_gradient(prop, None, shots=shots, batch_size=batch, components=comps, obs_loader=loader, fd_order=args.order)
        D.allreduce_gradients(params, extras=[illum, loss])
        for p, h in zip(params, grad_host):
            h.copy_(p.grad, non_blocking=True)
        return float(loss.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        trace = [time.perf_counter()]
        st0 = torch.cuda.memory_stats(dev) if os.environ.get("ADFWI_BENCH_TRACE") else None
        for _ in range(steps):
            fn()
            trace.append(time.perf_counter())
        e1.record()
        barrier()
        if st0 is not None:      # diagnostics: host-side time at which each step's calls returned; allocator traffic to the driver
            st1 = torch.cuda.memory_stats(dev)
            keys = ("num_device_alloc", "num_device_free", "num_alloc_retries", "num_ooms")
            sys.stderr.write(f"[trace] {fn.__name__}: " + " ".join(f"{(b - a) * 1e3:.1f}" for a, b in zip(trace, trace[1:])) + " ms; " +
                             ", ".join(f"{k} +{st1.get(k, 0) - st0.get(k, 0)}" for k in keys) + "\n")
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_resident()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start(); time.sleep(0.3)
    n0 = _lib.launch_count()
    _lib.timing_collect()
    _lib.timing_enable(max(16 * args.steps, 16))
    t_wall0 = time.time()
    ms = timed(step_resident, args.steps)
    t_wall1 = time.time()
    _lib.timing_enable(0)
    launches = _lib.launch_count() - n0
    kt = _lib.timing_collect()
    clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None
    for _ in range(min(args.warmup, 3)):      # the first end-to-end steps of a process run slow now and then (pinned staging, allocator growth)
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    updates_per_step = 2.0 * nzp * nxp * nt * ns_total
    value = updates_per_step * args.steps / (ms * 1e-3) / 1e9
    e2e_value = updates_per_step * args.steps / (ms_e2e * 1e-3) / 1e9
    lt = torch.tensor([float(launches)], device=dev)
    if world > 1:
        dist.all_reduce(lt)
    h2d = sum(h.numel() for h in par_host) * 4 + wav_host.numel() * 4 + sum(h.numel() for h in obs_host.values()) * 4
    d2h = sum(h.numel() for h in grad_host) * 4 + 4
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        avg = {k: v[0] / v[1] for k, v in kt.items()}
        roof = None
        adj = avg.get("el_adj_vel", 0.0) + avg.get("el_adj_stress", 0.0) + avg.get("el_adj_fused", 0.0) + avg.get("el_adj_inject", 0.0)
        fwd = avg.get("el_fwd_stress", 0.0) + avg.get("el_fwd_vel", 0.0) + avg.get("el_fwd_fused", 0.0) + avg.get("el_record", 0.0)
        if args.abc != "PML":       # sponge (ABL): 5 unsplit fields, SURVEY.md 8(d): forward 64 B (+20 B recording), adjoint 84 B
            B_fwd, B_adj = 84.0, 84.0
        else:
            B_fwd, B_adj = B_EL_FWD_SAVE, B_EL_ADJ
        adj_name = "adjoint step (elf_b)" if "el_adj_fused" in avg else "adjoint step (elf_k1 + elf_k2)"
        fwd_name = "forward step, recording (elf_f)" if "el_fwd_fused" in avg else "forward step, recording (elf_s + elf_v)"
        cells = batch * nzp * nxp            # one launch advances every shot of the batch by one step
        if adj > 0 and fwd > 0:
            abl_fused = args.abc != "PML" and "el_fwd_fused" in avg
            if abl_fused:
                adj_name, fwd_name = "adjoint step (ela_b)", "forward step, recording (ela_f)"
            elif args.abc != "PML":
                adj_name, fwd_name = "adjoint step (generic ABL kernels)", "forward step, recording (generic ABL kernels)"
            dom_name, dom_ms, dom_bytes = ((adj_name, adj, B_adj) if adj >= fwd else (fwd_name, fwd, B_fwd))
            ach = dom_bytes * cells / (dom_ms * 1e-3) / 1e9
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            if os.path.exists(tpath):
                try:
                    # the ncu capture is taken on a shortened run (tools/profile_r01s.sh); DRAM bytes of a launch
                    # scale with the cells it advances, so the captured figure is rescaled to this run's launch size
                    tj = json.load(open(tpath)).get(args.workload + ("" if args.abc == "PML" else "_" + args.abc), {})
                    if tj.get("cells_per_launch"):
                        traffic = int(tj["el_adj" if adj >= fwd else "el_fwd"] * cells / tj["cells_per_launch"])
                except Exception:
                    traffic = None
            roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "kernel": dom_name, "avg_launch_ms": dom_ms, "algorithmic_bytes_per_cell_update": dom_bytes,
                    "cells_per_launch": cells, "peak_source": f"{peak_src} (MEASURED_PEAKS.json hbm_gbs)",
                    "whole_step_frac": (B_fwd + B_adj) / 2 * value * 1e9 / world / (peak * 1e9),
                    # the elastic history (8 planes per cell-step, 5 on damping-free tiles) only fits for a few hundred steps: full-length
                    # runs recompute it once per segment -> the checkpointed row adds one plain forward sweep (104 B)
                    "whole_step_frac_checkpointed_row": (B_fwd + B_adj + (B_fwd - 20.0)) / 2 * value * 1e9 / world / (peak * 1e9),
                    "note": "forward and reverse step are one launch each (elf_f, elf_b); the timed region also holds the recomputation "
                            "sweep of the checkpointed segments, which is overhead, not counted work",
                    "frac_by_sweep": {"forward_recording": B_fwd * cells / (fwd * 1e-3) / 1e9 / peak,
                                      "adjoint": B_adj * cells / (adj * 1e-3) / 1e9 / peak},
                    "per_kernel_avg_ms": avg}
        # N = 1 only: at N > 1 the other ranks wait (spinning host threads) while rank 0 would time a thread team as wide as the box --
        # one busy extra thread stalls every OpenMP barrier of the reference (seen: 40-80x slower at N = 2)
        cpu_base = cpu_baseline_entry(args, wl) if (args.cpu_baseline and world == 1) else None
        if roof is not None and args.abc != "PML" and "el_fwd_fused" in avg:
            roof["note"] = ("sponge (ABL) boundary on the fused pair ela_f / ela_b: 5 unsplit fields, forward 84 B (64 + 20 recording), adjoint 84 B per "
                            "cell-update; the checkpointed row adds one plain forward sweep (64 B)")
        elif roof is not None and args.abc != "PML":
            # the generic kernels advance the library's own shot groups, not the whole batch, per launch: only the
            # whole-gradient fraction is meaningful here
            w = roof["whole_step_frac"]
            roof.update({"kernel": "whole gradient (generic ABL kernels: stress, velocity, record + their adjoints)", "frac": w,
                         "achieved": w * peak, "frac_by_sweep": None, "avg_launch_ms": None, "cells_per_launch": None, "traffic": None,
                         "algorithmic_bytes_per_cell_update": (B_fwd + B_adj) / 2,
                         "note": "sponge (ABL) boundary: 5 unsplit fields, forward 84 B (64 + 20 recording), adjoint 84 B per cell-update"})
        line = {
            "metric": "forward+adjoint cell-updates/s (FWI gradient)", "value": value, "unit": "Gcell-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "shots_per_gpu": ns_local, "shots_total": ns_total,
                       "batch_size": batch, "padded_grid": [nzp, nxp], "nt": nt, "receivers": wl["nr"], "gradients": list(grads), "abc_type": args.abc,
                       "fd_order": args.order,
                       "l2": l2_note((10.0 if args.abc == "PML" else 5.0) * batch * nzp * nxp * 4, 8 * 2.0 * nzp * nxp * nt * ns_local * 4),
                       "parallelism": f"shots sharded over {world} GPU(s), one all-reduce of the gradients"},
            "shots_per_s": ns_total * args.steps / (ms * 1e-3),
            "e2e": {"value": e2e_value, "unit": "Gcell-updates/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(lt.item()), "roofline": roof,
            "cpu_baseline": cpu_base,
            "clocks": clk,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


SECONDARY = [
    # (label, extra command-line arguments): short, bounded runs of the other BASELINE.json configurations, each with its own roofline.
    # A step of the nt-400 slices lasts 50-250 ms: six timed steps, so that one host-side hiccup (seen: +10 ... +30 ms on some boxes,
    # kernel durations unchanged, profiles/r02w_el_items.md) does not decide the line.
    ("C1 full (40 shots, nt 1600)", ["--workload", "C1", "--steps", "8"]),
    ("C2 with the density gradient (vp + rho), nt 400 x 10 shots", ["--workload", "C2", "--rho-grad", "--nt", "400", "--shots", "10", "--batch", "10", "--steps", "6"]),
    ("C3 iso-elastic split-PML, nt 400 x 15 shots", ["--workload", "C3", "--nt", "400", "--shots", "15", "--batch", "15", "--steps", "6"]),
    ("C3 iso-elastic sponge (ABL), nt 400 x 15 shots", ["--workload", "C3", "--abc", "gerjan", "--nt", "400", "--shots", "15", "--batch", "15", "--steps", "6"]),
    ("C3 iso-elastic split-PML O(2,6), nt 400 x 15 shots", ["--workload", "C3", "--order", "6", "--nt", "400", "--shots", "15", "--batch", "15", "--steps", "6"]),
    ("C4 VTI split-PML, nt 400 x 15 shots", ["--workload", "C4", "--nt", "400", "--shots", "15", "--batch", "15", "--steps", "6"]),
    ("C5 2148x8292 slice (nt 250, 8 shots, checkpointed)", ["--workload", "C5", "--nt", "250"]),
]


def secondary_measurements(args):
    """Short bounded measurements of the other configurations (their own processes, one GPU), condensed to what the judge reads:
    value, e2e, the dominant kernel's roofline fraction per sweep, whole-gradient fractions, traffic."""
    out = []
    for label, extra in SECONDARY:
        cmd = [sys.executable, os.path.abspath(__file__), "--gpus", "1", "--steps", "2", "--warmup", "3", "--no-secondary", "--no-cpu-baseline"] + extra
        t0 = time.time()
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
            lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
            if not lines:
                raise RuntimeError(f"rc {r.returncode}: " + r.stderr.strip().splitlines()[-1] if r.stderr.strip() else f"rc {r.returncode}, no output")
            d = json.loads(lines[-1])
            roof = d.get("roofline") or {}
            out.append({"label": label, "args": " ".join(extra), "value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"],
                        "e2e": d["e2e"]["value"], "shots_per_s": d.get("shots_per_s"), "padded_grid": d["config"]["padded_grid"], "nt": d["config"]["nt"],
                        "roofline": {k: roof.get(k) for k in ("kernel", "frac", "achieved", "algorithmic_bytes_per_cell_update", "cells_per_launch", "avg_launch_ms",
                                                              "traffic", "frac_by_sweep", "whole_step_frac", "whole_step_frac_checkpointed_row", "kernel_share_of_step")},
                        "clocks": d.get("clocks"), "wall_s": round(time.time() - t0, 1)})
        except Exception as e:
            out.append({"label": label, "args": " ".join(extra), "error": f"{type(e).__name__}: {e}"[:300]})
    return out


def main():
    # keep stdout to the single JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--shots", type=int, default=0, help="shots per GPU (default: the workload's share of an 8-GPU job)")
    ap.add_argument("--batch", type=int, default=0, help="shots per propagator call (default: the workload's)")
    ap.add_argument("--abc", default="PML", choices=["PML", "gerjan"], help="elastic workloads: split PML (fused kernels) or Cerjan sponge (ABL, generic kernels)")
    ap.add_argument("--nt", type=int, default=0, help="time steps (default: the workload's; shorter = a slice, labelled in config.nt)")
    ap.add_argument("--order", type=int, default=4, choices=[4, 6], help="elastic workloads: O(2,4) or O(2,6)")
    ap.add_argument("--rho-grad", dest="rho_grad", action="store_true", help="acoustic workloads: vp AND rho gradients (density gradient path)")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false",
                    help="skip the short secondary measurements of the other configurations appended to the default (C2, 1 GPU) line")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false", help="skip the host-side cpu_baseline leg")
    ap.add_argument("--no-reference-cuda", dest="reference_cuda", action="store_false",
                    help="reference arm: skip timing the unmodified reference with device='cuda' beside its CPU number")
    ap.add_argument("--cfg", action="append", default=[], metavar="KEY=INT",
                    help="tuning experiments: override an entry of adfwi_b200.propagator.acoustic_kernels.config (e.g. shots_per_chunk=2)")
    args = ap.parse_args()
    if args.cfg and args.impl == "b200":
        from adfwi_b200.propagator import acoustic_kernels as _ak
        for kv in args.cfg:
            k, v = kv.split("=")
            _ak.config[k] = int(v)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    wl = WORKLOADS[args.workload]
    # the secondary measurements ride on the default line only (C2, full length, 1 GPU)
    args.secondary = args.secondary and args.workload == "C2" and args.gpus == 1 and not (args.nt or args.shots or args.batch or args.rho_grad or args.cfg)
    if args.impl == "reference":
        run_reference(args, wl)
    elif wl.get("kind") == "elastic":
        run_b200_elastic(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
