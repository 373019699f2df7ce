"""Quick kernel-throughput probe (developer tool, not the contract bench).
usage: python tools/quick_bench.py [nz nx ns nt [G ...]]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adfwi_b200.propagator import acoustic_kernels as ak

def run(nz, nx, ns, nt, G, K=None, nabc=50, grad=True, nr=None):
    dev = torch.device("cuda:0")
    ak.config["shots_per_group"] = G
    ak.config["ckpt_interval"] = K
    nr = nr or nx
    v = (2000 + 1000 * torch.rand(nz, nx, device=dev)).requires_grad_(grad)
    rho = torch.full((nz, nx), 2000.0, device=dev)
    damp = torch.zeros(nz + 2 * nabc, nx + 2 * nabc, device=dev)
    sx = torch.linspace(2, nx - 3, ns, device=dev).long(); sz = torch.ones(ns, device=dev).long()
    rx = torch.arange(nr, device=dev).long() % nx; rz = torch.ones(nr, device=dev).long()
    src_v = torch.randn(ns, nt, device=dev)
    def once():
        if v.grad is not None: v.grad = None
        rec = ak.forward_kernel(nx, nz, 10.0, 10.0, nt, 1e-3, nabc, True, sx, sz, ns, src_v, rx, rz, nr, damp, v, rho, device=dev)
        if grad:
            (rec["p"] * rec["p"]).sum().backward()
    once(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); once(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    cells = (nz + 2 * nabc) * (nx + 2 * nabc) * ns * nt * (2 if grad else 1)
    print(f"nz={nz} nx={nx} ns={ns} nt={nt} G={G} K={K} grad={grad}: {ms:.1f} ms  {cells / ms / 1e6:.1f} Gcell-upd/s  mem={torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)

if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    nz, nx, ns, nt = (a + [350, 1700, 8, 200])[:4] if len(a) < 4 else a[:4]
    Gs = a[4:] or [0]
    for G in Gs:
        run(nz, nx, ns, nt, G, grad=False)
        run(nz, nx, ns, nt, G, grad=True)
