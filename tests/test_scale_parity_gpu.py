"""Parity at the BASELINE.json configurations themselves (round-1 verdict: every fixture carrying reference truth
was <= 64 x 88 cells and nt <= 260, where the fused kernels run 1-4 tiles and the adjoint round-off has not grown yet).

  * CUDA path vs the CPU oracle on the C1 grid IN FULL (148 x 260 padded, nt 1600, 4 shots) and on 2-shot x 400-step
    slices of the C2 (450 x 1800), C3 (402 x 1800, iso-elastic split-PML) and C4 (372 x 820, VTI) grids and a 1-shot x 40-step
    slice of the C5 grid (2148 x 8292): records
    bit-identical, coefficient-plane gradients and model-level gradients (vp, rho / vp, vs, rho / eps, delta) within the
    1e-4 bar of BASELINE.json, with a realistic cotangent (L2 waveform misfit against records of a "true" model);
  * CUDA path vs committed fixtures of the UNMODIFIED reference at its examples' own grid size and nt
    (tests/golden/make_golden_scale.py): acoustic Marmousi2 example (nt 1600) and the VTI example (nt 1000), through
    the propagator API + L2 misfit + backward.

Everything goes through the drop-in forward_kernel / Propagator -> C ABI.  The oracle (oracle/) is only the checker."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REC_TOL = 1e-5     # BASELINE.json: records rel-L2 <= 1e-5
GRAD_TOL = 1e-4    # BASELINE.json: gradients rel-L2 <= 1e-4


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def l2_misfit(syn_rec, obs, dt):
    """Misfit_waveform_L2 (fwi/misfit/L2.py:22-28).  The tiny offset keeps the derivative of the square root finite for traces
    the wave has not reached yet (zero residual -> 0/0 = NaN in the reference's expression); the resulting cotangent is handed
    to the CUDA path and to the oracle alike."""
    r = obs - syn_rec
    return torch.sum(torch.sqrt(torch.sum(r * r * dt, dim=1) + 1e-30))


def _oracle_threads():
    from oracle import oracle as O
    O.lib().oracle_set_threads(os.cpu_count() or 1)
    return O


# ---------------------------------------------------------------------------------------------------
# acoustic: C1 in full, C2 slice
# ---------------------------------------------------------------------------------------------------
def _acoustic_vs_oracle(nz, nx, nabc, nt, dx, dt, f0, ns, nr, tag):
    from adfwi_b200 import synthetic as syn
    from adfwi_b200.propagator import acoustic_kernels as ak
    from adfwi_b200.propagator.boundary_condition import bc_pml
    O = _oracle_threads()
    dev = torch.device("cuda:0")
    vp_true = syn.marmousi_like_vp(nz, nx)
    vp0 = syn.smooth2d(vp_true, 6)
    rho0 = syn.gardner_rho(vp0)
    damp = bc_pml(nx, nz, dx, dx, pml=nabc, vmax=float(vp_true.max()), free_surface=False).astype(np.float32)
    sx = np.round(np.linspace(2, nx - 3, ns)).astype(np.int64); sz = np.ones(ns, np.int64)
    rx = np.round(np.linspace(0, nx - 1, nr)).astype(np.int64); rz = np.ones(nr, np.int64)
    wav = np.broadcast_to(syn.integrated_ricker(nt, dt, f0).astype(np.float32), (ns, nt)).copy()
    t = lambda a: torch.tensor(a, device=dev)
    run = lambda v, r: ak.forward_kernel(nx, nz, dx, dx, nt, dt, nabc, True, t(sx), t(sz), ns, t(wav), t(rx), t(rz), nr, t(damp), v, r,
                                         checkpoint_segments=1, device=dev)
    with torch.no_grad():
        obs = run(t(vp_true), t(syn.gardner_rho(vp_true)))["p"]
    out = {}
    for mode in ("vp", "vp_rho"):       # vp only (the configuration of every reference example) and vp + rho
        v = t(vp0).requires_grad_(True)
        r = t(rho0).requires_grad_(mode == "vp_rho")
        rec = run(v, r)
        loss = l2_misfit(rec["p"], obs, dt)
        g_rec, = torch.autograd.grad(loss, rec["p"], retain_graph=True)
        loss.backward()
        out[mode] = (rec, g_rec, v.grad.cpu().numpy(), None if r.grad is None else r.grad.cpu().numpy())
    rec, g_rec = out["vp_rho"][0], out["vp_rho"][1]
    assert torch.equal(out["vp"][0]["p"], rec["p"])
    coef = O.acoustic_coefficients(vp0, rho0, damp, dt, dx, nabc, True)
    ref = O.acoustic_run(coef, nabc, True, dt, sx, sz, wav, rx, rz, g_rcv=(g_rec.cpu().numpy(), None, None), need_g_alpha2=True)
    for k in "puw":
        got = rec[k].detach().cpu().numpy()
        assert rel_l2(got, ref[k]) <= REC_TOL, (tag, k)
        assert np.array_equal(got, ref[k]), f"{tag}: record {k} not bit-identical to the oracle"
    gv, grho = O.acoustic_model_gradients(coef, ref["g_alpha1"], ref["g_alpha2"], dt, dx, nabc)
    gv_only, _ = O.acoustic_model_gradients(coef, ref["g_alpha1"], None, dt, dx, nabc)
    # with rho detached the vp gradient only sees alpha1's dependence on c -- same expression in both cases
    errs = dict(g_vp=rel_l2(out["vp"][2], gv_only), g_vp_with_rho=rel_l2(out["vp_rho"][2], gv), g_rho=rel_l2(out["vp_rho"][3], grho))
    print(f"{tag}: gradient rel-L2 vs oracle {errs}")
    assert max(errs.values()) <= GRAD_TOL, (tag, errs)
    return errs


def test_c1_full_vs_oracle():
    """C1 = iso-acoustic Marmousi2 example grid 88 x 200 (+30 -> 148 x 260), dx 40 m, dt 3 ms, nt 1600, f0 5 Hz."""
    _acoustic_vs_oracle(88, 200, 30, 1600, 40.0, 3e-3, 5.0, ns=4, nr=200, tag="C1 full (nt 1600, 4 shots)")


def test_c2_slice_vs_oracle():
    """C2 = 350 x 1700 (+50 -> 450 x 1800), dx 10 m, dt 1 ms: 2 shots x 400 steps (f0 raised so that the wave has left the source)."""
    _acoustic_vs_oracle(350, 1700, 50, 400, 10.0, 1e-3, 25.0, ns=2, nr=1700, tag="C2 grid slice (nt 400, 2 shots)")


def test_c5_slice_vs_oracle():
    """C5 = 2048 x 8192 (+50 -> 2148 x 8292, 4422 tiles of 64 x 32), dx 5 m, dt 0.5 ms: 1 shot x 40 steps (the oracle keeps its
    history in host memory: 8.5 GB for this slice)."""
    import psutil
    if psutil.virtual_memory().available < 40 * 2 ** 30:
        pytest.skip("needs 40 GB of free host memory for the oracle's history")
    _acoustic_vs_oracle(2048, 8192, 50, 40, 5.0, 5e-4, 100.0, ns=1, nr=2048, tag="C5 grid slice (nt 40, 1 shot)")


# ---------------------------------------------------------------------------------------------------
# elastic: C3 (iso) and C4 (VTI) grid slices
# ---------------------------------------------------------------------------------------------------
PLANES = ("C11", "C13", "C33", "C55", "bx", "bz")
COMPS = ("txx", "tzz", "txz", "vx", "vz")
COEF_IDX = {"C11": 0, "C13": 2, "C33": 11, "C55": 18}


def _elastic_vs_oracle(nz, nx, nabc, nt, dx, dt, f0, ns, nr, z_sr, vti, params, tag, abc="PML", order=4):
    from adfwi_b200 import synthetic as syn
    from adfwi_b200.propagator import ElasticPropagator, elastic_kernels as ek
    O = _oracle_threads()
    dev = torch.device("cuda:0")
    vp_true = syn.marmousi_like_vp(nz, nx)
    vp0 = syn.smooth2d(vp_true, 6)
    mk_vs = lambda v: (v / np.sqrt(3.0)).astype(np.float32)
    eps = np.full((nz, nx), 0.1 if vti else 0.0, np.float32); delta = np.full((nz, nx), -0.1 if vti else 0.0, np.float32)
    eps_true = eps.copy()
    if vti:
        eps_true[nz // 3:nz // 2, nx // 3:nx // 2] = 0.2
    survey = syn.surface_survey(nx, ns, nr, nt, dt, f0, src_z=z_sr, rcv_z=z_sr)
    mk = lambda v, e, req, d: syn.ElasticGridModel(v, mk_vs(v), syn.gardner_rho(v), eps=e, delta=delta, dx=dx, dz=dx, nabc=nabc,
                                                   free_surface=True, abc_type=abc, requires_grad=req, device=d)
    true_model, model = mk(vp_true, eps_true, (), dev), mk(vp0, eps, params, dev)
    prop_true = ElasticPropagator(true_model, survey, device=dev)
    prop = ElasticPropagator(model, survey, device=dev)
    pml = abc == "PML"
    if pml:
        prop.bcx, prop.bcz = prop_true.bcx, prop_true.bcz
    else:
        prop.damp = prop_true.damp
    with torch.no_grad():
        o = prop_true.forward(fd_order=order)
        obs = {c: o[c] for c in ("vx", "vz")}
    # model-level gradients through the propagator + parameterisation
    rec = prop.forward(fd_order=order)
    loss = l2_misfit(rec["vx"], obs["vx"], dt) + l2_misfit(rec["vz"], obs["vz"], dt)
    g_vx, g_vz = torch.autograd.grad(loss, [rec["vx"], rec["vz"]], retain_graph=True)
    loss.backward()
    g_model = {k: getattr(model, k).grad.cpu().numpy() for k in params}
    # kernel-level gradients: the six planes as leaves, same cotangent
    model.forward()
    L = {k: model.CC[i].detach().clone().requires_grad_(True) for k, i in COEF_IDX.items()}
    L["bx"] = model.bx.detach().clone().requires_grad_(True); L["bz"] = model.bz.detach().clone().requires_grad_(True)
    CC = list(model.CC)
    for k, i in COEF_IDX.items():
        CC[i] = L[k]
    rec2 = ek.forward_kernel(nx, nz, dx, dx, nt, dt, nabc, True, prop.src_x, prop.src_z, ns, prop.wavelet, prop.moment_tensor,
                             prop.rcv_x, prop.rcv_z, nr, abc, prop.bcx if pml else None, prop.bcz if pml else None, None if pml else prop.damp,
                             None, None, L["bx"], L["bz"], CC, fd_order=order, n_segments=1, device=dev)
    for c in COMPS:
        assert torch.equal(rec2[c], rec[c]), c
    ((rec2["vx"] * g_vx).sum() + (rec2["vz"] * g_vz).sum()).backward()
    # oracle on the same planes and the same cotangent
    planes = {k: L[k].detach().cpu().numpy() for k in PLANES}
    src = survey.source
    bc = dict(bcx=prop.bcx.cpu().numpy(), bcz=prop.bcz.cpu().numpy()) if pml else dict(damp=prop.damp.cpu().numpy())
    ref = O.elastic_run(planes, abc, order, True, nz, nx, nabc, dx, dx, dt, src.loc[:, 0], src.loc[:, 1], src.wavelet, src.moment_tensor,
                        survey.receiver.loc[:, 0], survey.receiver.loc[:, 1], g_rcv=[None, None, None, g_vx.cpu().numpy(), g_vz.cpu().numpy()], **bc)
    for c in COMPS:
        got = rec[c].detach().cpu().numpy()
        assert rel_l2(got, ref[c]) <= REC_TOL, (tag, c)
        assert np.array_equal(got, ref[c]), f"{tag}: record {c} not bit-identical to the oracle"
    errs = {k: rel_l2(L[k].grad.cpu().numpy(), ref["g_own"][k]) for k in PLANES}
    # chain the oracle's plane gradients through the parameterisation (plain torch on the CPU) to the model parameters
    cpu_model = mk(vp0, eps, params, "cpu")
    cpu_model.forward()
    outs = [cpu_model.CC[COEF_IDX[k]] if k in COEF_IDX else getattr(cpu_model, k) for k in PLANES]
    gm = torch.autograd.grad(outs, [getattr(cpu_model, k) for k in params], grad_outputs=[torch.tensor(ref["g_own"][k], dtype=torch.float32) for k in PLANES])
    for k, gref in zip(params, gm):
        errs["model_" + k] = rel_l2(g_model[k], gref.numpy())
    print(f"{tag}: gradient rel-L2 vs oracle {errs}")
    assert max(errs.values()) <= GRAD_TOL, (tag, errs)
    return errs


def test_c3_slice_vs_oracle():
    """C3 = iso-elastic 350 x 1700 (+50, free surface -> 402 x 1800), split-PML O(2,4): 2 shots x 400 steps, vp / vs / rho."""
    _elastic_vs_oracle(350, 1700, 50, 400, 10.0, 1e-3, 25.0, ns=2, nr=1700, z_sr=10, vti=False, params=("vp", "vs", "rho"),
                       tag="C3 grid slice (nt 400, 2 shots)")


def test_c3_sponge_slice_vs_oracle():
    """The C3 grid with the multiplicative sponge (ABL) boundary: the fused pair ela_f / ela_b, 2 shots x 400 steps."""
    _elastic_vs_oracle(350, 1700, 50, 400, 10.0, 1e-3, 25.0, ns=2, nr=1700, z_sr=10, vti=False, params=("vp", "vs", "rho"),
                       tag="C3 grid slice, sponge boundary (nt 400, 2 shots)", abc="gerjan")


def test_c3_o26_slice_vs_oracle():
    """The C3 grid at O(2,6): elf_f<3> forward, the elf_k1 + elf_k2 pair in reverse, 2 shots x 300 steps."""
    _elastic_vs_oracle(350, 1700, 50, 300, 10.0, 1e-3, 25.0, ns=2, nr=1700, z_sr=10, vti=False, params=("vp", "vs", "rho"),
                       tag="C3 grid slice, O(2,6) (nt 300, 2 shots)", order=6)


def test_c4_slice_vs_oracle():
    """C4 scale-up = VTI 320 x 720 (-> 372 x 820), dx 2.5 m, dt 0.25 ms: 2 shots x 400 steps, eps / delta (+ vp, vs, rho)."""
    _elastic_vs_oracle(320, 720, 50, 400, 2.5, 2.5e-4, 80.0, ns=2, nr=720, z_sr=10, vti=True, params=("eps", "delta", "vp", "vs", "rho"),
                       tag="C4 grid slice (nt 400, 2 shots)")


# ---------------------------------------------------------------------------------------------------
# unmodified reference at example scale (committed fixtures)
# ---------------------------------------------------------------------------------------------------
def test_reference_acoustic_example_scale(golden_dir):
    """Acoustic Marmousi2 example geometry, nt 1600, 4 shots: AcousticPropagator + L2 misfit + backward against the
    UNMODIFIED reference run on CPU (tests/golden/make_golden_scale.py)."""
    from adfwi_b200 import synthetic as syn
    from adfwi_b200.propagator import AcousticPropagator
    g = np.load(f"{golden_dir}/acoustic_c1_scale.npz")
    dev = torch.device("cuda:0")
    nt, dt, f0 = int(g["nt"]), float(g["dt"]), float(g["f0"])
    # rho as the reference model derived it (numpy power on the host); auto_update_rho would redo it with torch's pow on the device
    model = syn.AcousticGridModel(g["vp_init"], rho=g["rho_init"], dx=float(g["dx"]), dz=float(g["dz"]), nabc=int(g["nabc"]),
                                  free_surface=True, vp_grad=True, auto_update_rho=False, device=dev)
    src = syn.Source(np.stack([g["src_x"], g["src_z"]], 1), g["wavelet"], nt, dt, f0)
    rcv = syn.Receiver(np.stack([g["rcv_x"], g["rcv_z"]], 1))
    prop = AcousticPropagator(model, syn.Survey(src, rcv), device=dev)
    prop.damp = torch.tensor(g["damp"], device=dev)
    rec = prop.forward(checkpoint_segments=4)
    loss = l2_misfit(rec["p"], torch.tensor(g["obs_p"], device=dev), dt)
    loss.backward()
    e_rec = {k: rel_l2(rec[k].detach().cpu().numpy()[:, :, ::s], g["rec_" + k]) for k, s in (("p", 1), ("u", 5), ("w", 5))}
    e_loss = abs(float(loss) - float(g["loss"])) / float(g["loss"])
    e_g = rel_l2(model.vp.grad.cpu().numpy(), g["g_vp"])
    e_ill = rel_l2(rec["forward_wavefield_p"].cpu().numpy(), g["illum_p"])
    print(f"reference, acoustic example scale (nt {nt}): records {e_rec}, loss {e_loss:.2e}, g_vp {e_g:.2e}, illumination {e_ill:.2e}")
    assert max(e_rec.values()) <= REC_TOL
    assert np.array_equal(rec["p"].detach().cpu().numpy(), g["rec_p"]), "records not bit-identical to the reference at nt 1600"
    assert e_loss <= 1e-5 and e_g <= GRAD_TOL and e_ill <= 1e-4


def test_reference_vti_example_scale(golden_dir):
    """VTI example (80 x 180, nabc 50, nt 1000, sources at z = 70, receivers at z = 10), 2 shots: ElasticPropagator + L2
    misfit on vx, vz + backward against the UNMODIFIED reference (AnisotropicElasticModel parameterisation) on CPU."""
    from adfwi_b200 import synthetic as syn
    from adfwi_b200.propagator import ElasticPropagator
    g = np.load(f"{golden_dir}/vti_c4_scale.npz")
    dev = torch.device("cuda:0")
    nt, dt, f0 = int(g["nt"]), float(g["dt"]), float(g["f0"])
    params = ("vp", "vs", "rho", "eps", "delta")
    model = syn.ElasticGridModel(g["vp"], g["vs"], g["rho"], eps=g["eps_init"], delta=g["delta"], dx=float(g["dx"]), dz=float(g["dz"]),
                                 nabc=int(g["nabc"]), free_surface=True, abc_type="PML", requires_grad=params, device=dev)
    src = syn.Source(np.stack([g["src_x"], g["src_z"]], 1), g["wavelet"], nt, dt, f0)
    rcv = syn.Receiver(np.stack([g["rcv_x"], g["rcv_z"]], 1))
    prop = ElasticPropagator(model, syn.Survey(src, rcv), device=dev)
    prop.bcx, prop.bcz = torch.tensor(g["bcx"], device=dev), torch.tensor(g["bcz"], device=dev)
    rec = prop.forward(fd_order=4, checkpoint_segments=4)
    loss = l2_misfit(rec["vx"], torch.tensor(g["obs_vx"], device=dev), dt) + l2_misfit(rec["vz"], torch.tensor(g["obs_vz"], device=dev), dt)
    loss.backward()
    e_rec = {k: rel_l2(rec[k].detach().cpu().numpy()[:, :, ::s], g["rec_" + k]) for k, s in (("vx", 1), ("vz", 1), ("txx", 6), ("tzz", 6), ("txz", 6))}
    e_loss = abs(float(loss) - float(g["loss"])) / float(g["loss"])
    e_g = {k: rel_l2(getattr(model, k).grad.cpu().numpy(), g["g_" + k]) for k in params}
    print(f"reference, VTI example scale (nt {nt}): records {e_rec}, loss {e_loss:.2e}, gradients {e_g}")
    assert max(e_rec.values()) <= REC_TOL
    assert e_loss <= 1e-5
    assert max(e_g.values()) <= GRAD_TOL, e_g
