"""Iteration-level parity: three iterations of the reference's AcousticFWI loop (CPU fixture of the unmodified
reference, tests/golden/make_golden_fwi.py: shot batches of 2, per-trace normalisation, L2 misfit, GradProcessor
"Marine" mute + illumination preconditioner + smoothing + max-normalisation, SGD + StepLR) against the same loop
run on the device through AcousticPropagator (C ABI) + the device-side GradProcessor + torch's own optimiser."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_three_fwi_iterations_match_reference(golden_dir):
    import fwi_loops as fwi
    from adfwi_b200 import synthetic as syn
    from adfwi_b200.propagator import AcousticPropagator, GradProcessor
    g = np.load(f"{golden_dir}/fwi_acoustic_3iter.npz")
    dev = torch.device("cuda:0")
    nt, dt, f0 = int(g["nt"]), float(g["dt"]), float(g["f0"])
    model = syn.AcousticGridModel(g["vp_init"], dx=float(g["dx"]), dz=float(g["dz"]), nabc=int(g["nabc"]), free_surface=True,
                                  vp_grad=True, device=dev)
    src = syn.Source(np.stack([g["src_x"], g["src_z"]], 1), g["wavelet"], nt, dt, f0)
    rcv = syn.Receiver(np.stack([g["rcv_x"], g["rcv_z"]], 1))
    prop = AcousticPropagator(model, syn.Survey(src, rcv), device=dev)
    prop.damp = torch.tensor(g["damp"], device=dev)
    opt = torch.optim.SGD(model.parameters(), lr=float(g["lr"]))
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=int(g["step_size"]), gamma=float(g["gamma"]))
    gp = GradProcessor(grad_mute=int(g["grad_mute"]), grad_smooth=int(g["grad_smooth"]), norm_grad=True, forw_illumination=True,
                       marine_or_land="Marine")
    hist = fwi.acoustic_fwi(prop, model, opt, sched, torch.tensor(g["obs_p"], device=dev), iterations=3, batch_size=int(g["batch_size"]),
                            gradient_processor=gp, waveform_normalize=True)
    errs = dict(loss=[abs(a - b) / b for a, b in zip(hist["loss"], g["iter_loss"])],
                grad=[rel_l2(a.cpu().numpy(), b) for a, b in zip(hist["grad"], g["iter_grad"])],
                vp=float(np.abs(model.vp.detach().cpu().numpy() - g["iter_vp"][-1]).max()))
    print("fwi iteration parity:", errs)
    assert hist["loss"][2] < hist["loss"][1] < hist["loss"][0]
    assert max(errs["loss"]) < 1e-4                      # losses, relative
    assert max(errs["grad"]) < 1e-3                      # processed gradients (max-normalised to vmax), relative L2
    assert errs["vp"] < 0.05                             # m/s, against model updates of up to 50 m/s


def test_two_elastic_fwi_iterations_match_reference(golden_dir):
    """Same at the elastic level: ElasticFWI.forward of the unmodified reference (vx + vz misfit, vp / vs / rho updated,
    gradient processor per parameter, SGD + StepLR; tests/golden/make_golden_fwi_elastic.py) against the device loop
    through ElasticPropagator (fused split-PML kernels) and the torch parameterisation."""
    import fwi_loops as fwi
    from adfwi_b200 import synthetic as syn
    from adfwi_b200.propagator import ElasticPropagator, GradProcessor
    g = np.load(f"{golden_dir}/fwi_elastic_2iter.npz")
    dev = torch.device("cuda:0")
    nt, dt, f0 = int(g["nt"]), float(g["dt"]), float(g["f0"])
    model = syn.ElasticGridModel(g["vp_init"], g["vs_init"], g["rho_init"], dx=float(g["dx"]), dz=float(g["dz"]), nabc=int(g["nabc"]),
                                 free_surface=True, abc_type="PML", requires_grad=("vp", "vs", "rho"), device=dev)
    src = syn.Source(np.stack([g["src_x"], g["src_z"]], 1), g["wavelet"], nt, dt, f0)
    rcv = syn.Receiver(np.stack([g["rcv_x"], g["rcv_z"]], 1))
    prop = ElasticPropagator(model, syn.Survey(src, rcv), device=dev)
    prop.bcx, prop.bcz = torch.tensor(g["bcx"], device=dev), torch.tensor(g["bcz"], device=dev)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=float(g["lr"]))
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=int(g["step_size"]), gamma=float(g["gamma"]))
    gp = GradProcessor(grad_mute=int(g["grad_mute"]), grad_smooth=int(g["grad_smooth"]), norm_grad=True, forw_illumination=True,
                       marine_or_land="Marine")
    obs = {c: torch.tensor(g["obs_" + c], device=dev) for c in ("vx", "vz")}
    hist = fwi.elastic_fwi(prop, model, opt, sched, obs, iterations=2, batch_size=int(g["batch_size"]), gradient_processor=gp)
    errs = dict(loss=[abs(a - b) / b for a, b in zip(hist["loss"], g["iter_loss"])],
                grad={k: [rel_l2(a.cpu().numpy(), b) for a, b in zip(hist["grad"][k], g["iter_grad_" + k])] for k in ("vp", "vs", "rho")},
                model={k: float(np.abs(getattr(model, k).detach().cpu().numpy() - g["final_" + k]).max()) for k in ("vp", "vs", "rho")})
    print("elastic fwi iteration parity:", errs)
    assert hist["loss"][1] < hist["loss"][0]
    assert max(errs["loss"]) < 1e-4
    assert max(max(v) for v in errs["grad"].values()) < 1e-3
    assert max(errs["model"].values()) < 0.05            # m/s (kg/m^3 for rho), against updates of 25-40
