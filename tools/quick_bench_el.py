"""Quick elastic kernel-throughput probe (developer tool).
usage: python tools/quick_bench_el.py [nz nx ns nt [pml(1/0) fd_order fs(1/0)]]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adfwi_b200.propagator import elastic_kernels as ek

def run(nz, nx, ns, nt, pml=True, fd_order=4, fs=True, nabc=50, grad=True):
    dev = torch.device("cuda:0")
    NN = fd_order // 2
    vp = 2500 + 1000 * torch.rand(nz, nx, device=dev); vs = vp / 1.73; rho = torch.full((nz, nx), 2000.0, device=dev)
    C33 = (vp * vp * rho).requires_grad_(grad); C55f = vs * vs * rho
    C11 = (vp * vp * rho * 1.1).requires_grad_(grad); C13 = (C33.detach() - 2 * C55f).requires_grad_(grad)
    C55 = C55f[1:-1, 1:-1].clone().requires_grad_(grad)
    b = 1.0 / rho
    bx = (0.5 * (b[:, :-1] + b[:, 1:])).requires_grad_(grad); bz = (0.5 * (b[:-1] + b[1:])).requires_grad_(grad)
    CC = [None] * 21
    CC[0], CC[2], CC[11], CC[18] = C11, C13, C33, C55
    nzp = nz + (nabc + NN if fs else 2 * nabc + NN); nxp = nx + 2 * nabc
    bcx = torch.zeros(nzp - NN, nxp, device=dev); bcz = torch.zeros(nzp - NN, nxp, device=dev)
    damp = torch.ones(nzp - NN, nxp, device=dev)
    sx = torch.linspace(2, nx - 3, ns, device=dev).long(); sz = torch.full((ns,), 10, device=dev).long()
    nr = nx
    rx = torch.arange(nr, device=dev).long(); rz = torch.full((nr,), 10, device=dev).long()
    src_v = torch.randn(ns, nt, device=dev)
    MT = torch.eye(3, device=dev).repeat(ns, 1, 1)
    def once():
        for t in (C11, C13, C33, C55, bx, bz):
            t.grad = None
        rec = ek.forward_kernel(nx, nz, 10.0, 10.0, nt, 1e-3, nabc, fs, sx, sz, ns, src_v, MT, rx, rz, nr,
                                "PML" if pml else "gerjan", bcx, bcz, damp, None, None, bx, bz, CC, fd_order=fd_order, device=dev)
        if grad:
            ((rec["vx"] * rec["vx"]).sum() + (rec["vz"] * rec["vz"]).sum()).backward()
    once(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); once(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    cells = nzp * nxp * ns * nt * (2 if grad else 1)
    print(f"elastic nz={nz} nx={nx} ns={ns} nt={nt} pml={pml} O{fd_order} fs={fs} grad={grad}: {ms:.1f} ms  {cells / ms / 1e6:.1f} Gcell-upd/s  mem={torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)

if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    nz, nx, ns, nt = (a + [350, 1700, 4, 100])[:4] if len(a) < 4 else a[:4]
    pml = bool(a[4]) if len(a) > 4 else True
    fd = a[5] if len(a) > 5 else 4
    fs = bool(a[6]) if len(a) > 6 else True
    run(nz, nx, ns, nt, pml, fd, fs, grad=False)
    run(nz, nx, ns, nt, pml, fd, fs, grad=True)
