#!/usr/bin/env python
"""Small forward + adjoint cases of every fused pipeline, for `compute-sanitizer --tool memcheck|racecheck|synccheck`
(golden-sized grids so that the instrumented run finishes in a minute).  Prints the parity numbers it also checks."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from adfwi_b200.propagator import acoustic_kernels as ak, elastic_kernels as ek  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
dev = torch.device("cuda:0")


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def acoustic(persistent):
    g = np.load(os.path.join(G, "acoustic_fs.npz"))
    t = lambda k: torch.tensor(g[k], device=dev)
    ak.config["persistent"] = persistent
    v = t("vp").requires_grad_(True)
    rec = ak.forward_kernel(int(g["nx"]), int(g["nz"]), float(g["dx"]), float(g["dz"]), int(g["nt"]), float(g["dt"]), int(g["nabc"]), True,
                            t("src_x"), t("src_z"), len(g["src_x"]), t("src_v"), t("rcv_x"), t("rcv_z"), len(g["rcv_x"]), t("damp"), v, t("rho"), device=dev)
    (rec["p"] * t("W_p")).sum().backward()
    torch.cuda.synchronize()
    ok = np.array_equal(rec["p"].detach().cpu().numpy(), g["rec_p"])
    print(f"acoustic ({'acp_fwd/acp_adj' if persistent else 'ac_fwd_fused/ac_adj_fused'}): records identical {ok}, g_vp {rel(v.grad.cpu().numpy(), g['g_v_p']):.1e}")
    assert ok


def elastic(name, abc):
    g = np.load(os.path.join(G, name + ".npz"))
    t = lambda k: torch.tensor(g[k], device=dev)
    nz, nx = int(g["nz"]), int(g["nx"])
    L = {k: t("in_" + k).requires_grad_(True) for k in ("C11", "C13", "C33", "C55", "bx", "bz")}
    CC = [torch.zeros((nz, nx), device=dev)] * 21
    CC[0], CC[2], CC[11], CC[18] = L["C11"], L["C13"], L["C33"], L["C55"]
    pml = abc == "PML"
    rec = ek.forward_kernel(nx, nz, float(g["dx"]), float(g["dz"]), int(g["nt"]), float(g["dt"]), int(g["nabc"]), bool(g["free_surface"]),
                            t("src_x"), t("src_z"), len(g["src_x"]), t("src_v"), t("mt"), t("rcv_x"), t("rcv_z"), len(g["rcv_x"]), abc,
                            t("bcx") if pml else None, t("bcz") if pml else None, None if pml else t("damp"), None, None, L["bx"], L["bz"], CC,
                            fd_order=int(g["order"]), n_segments=int(g["segments"]), device=dev)
    sum((rec[k] * t("W_" + k)).sum() for k in ("vx", "vz")).backward()
    torch.cuda.synchronize()
    ok = np.array_equal(rec["vz"].detach().cpu().numpy(), g["rec_vz"])
    print(f"elastic {name}: records identical {ok}, g_C11 {rel(L['C11'].grad.cpu().numpy(), g['g_C11_vel']):.1e}")
    assert ok


def callers():
    """The 8(f) kernels: fused misfits, a regulariser, the Thomsen -> staggered-moduli chain with its fused padding, GradProcessor."""
    from adfwi_b200.fwi import misfit as M, regularization as Rg
    from adfwi_b200.model import thomsen_to_staggered_planes
    from adfwi_b200.propagator import GradProcessor
    g = np.load(os.path.join(G, "objective_misfit.npz"))
    for cls in (M.Misfit_waveform_L2, M.Misfit_global_correlation):
        sy = torch.tensor(g["syn"], device=dev, requires_grad=True)
        cls(dt=1.0, normalize=True).forward(torch.tensor(g["obs"], device=dev), sy).backward()
    rng = np.random.default_rng(0)
    m = torch.tensor((3000 + 500 * rng.random((60, 90))).astype(np.float32), device=dev, requires_grad=True)
    for cls in (Rg.TV_1order, Rg.Tikhonov_2order):
        cls(90, 60, 10.0, 10.0, 1e-3, 1e-3).forward(m).backward()
    e = np.load(os.path.join(G, "elastic_pml_o4_fs.npz"))
    par = [torch.tensor(e[k], device=dev, requires_grad=True) for k in ("vp", "vs", "rho", "eps", "delta")]
    planes = thomsen_to_staggered_planes(*par, anisotropic_type="vti")
    sum(p.sum() for p in planes).backward()
    gp = np.load(os.path.join(G, "gradproc_land_full.npz"))
    GradProcessor(grad_mute=int(gp["kw_grad_mute"]), grad_smooth=int(gp["kw_grad_smooth"]), norm_grad=True, forw_illumination=True,
                  marine_or_land="land").forward(nx=int(gp["nx"]), nz=int(gp["nz"]), vmax=gp["vmax"][()],
                                                 grad=torch.tensor(gp["grad"], device=dev), forw=torch.tensor(gp["forw"], device=dev))
    torch.cuda.synchronize()
    print("callers: misfits, regularisers, moduli + pad, gradient post-processing ran")


if __name__ == "__main__":
    which = sys.argv[1:] or ["acoustic", "persist", "pml", "abl", "callers"]
    if "acoustic" in which:
        acoustic(False)
    if "persist" in which:
        acoustic(True)
    if "pml" in which:
        elastic("elastic_pml_o4_fs", "PML")
    if "abl" in which:
        elastic("elastic_gerjan_o4_fs", "gerjan")
    if "callers" in which:
        callers()
