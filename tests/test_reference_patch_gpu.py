"""Boundary proof (test 5): the UNMODIFIED reference package with ``adfwi_b200.patch()`` applied.

The reference's own classes -- AcousticModel / IsotropicElasticModel, Survey, AcousticPropagator / ElasticPropagator,
AcousticFWI / ElasticFWI with their per-trace normalisation, Misfit_waveform_L2, host-side GradProcessor, torch optimiser and
scheduler -- run on the GPU with ``device="cuda"``; the only thing replaced is the module global ``forward_kernel`` of the two
propagator modules (acoustic_propagator.py:19,147 / elastic_propagator.py:16,132).  The result is compared with the
committed fixtures of the SAME scripts run unpatched on CPU (tests/golden/make_golden_fwi*.py).

Needs the reference tree: /root/reference in the build container, baseline/_ref (baseline/stage_reference.py) on the GPU box."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_loader

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="reference tree not staged (baseline/stage_reference.py)")]

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="module")
def patched():
    import adfwi_b200
    ref_loader.load()
    import ADFWI.propagator.acoustic_propagator as ap
    import ADFWI.propagator.elastic_propagator as ep
    saved = (ap.forward_kernel, ep.forward_kernel)
    mods = adfwi_b200.patch()
    assert sorted(mods) == ["ADFWI.propagator.acoustic_propagator", "ADFWI.propagator.elastic_propagator"]
    yield
    ap.forward_kernel, ep.forward_kernel = saved


def test_patched_reference_acoustic_fwi_matches_unpatched_cpu_run(patched, golden_dir):
    import make_golden_fwi as M
    from adfwi_b200 import _lib
    g = np.load(f"{golden_dir}/fwi_acoustic_3iter.npz")
    n0 = _lib.launch_count()
    fwi, model, prop, inp = M.build("cuda:0")
    assert np.array_equal(inp["obs_p"], g["obs_p"]), "observed records (reference forward through the patch) differ from the CPU run"
    fwi.forward(iteration=3, batch_size=2, checkpoint_segments=1)
    assert _lib.launch_count() > n0, "the patched reference did not reach libadfwi_b200.so"
    loss = np.array(fwi.iter_loss, dtype=np.float64)
    e_loss = np.abs(loss - g["iter_loss"]) / g["iter_loss"]
    e_grad = [rel_l2(a, b) for a, b in zip(fwi.iter_vp_grad, g["iter_grad"])]
    e_vp = float(np.abs(np.stack(fwi.iter_vp)[-1] - g["iter_vp"][-1]).max())
    print(f"patched reference AcousticFWI on cuda vs unpatched CPU: loss {e_loss}, processed gradients {e_grad}, final vp max diff {e_vp:.2e} m/s")
    assert e_loss.max() < 1e-4 and max(e_grad) < 1e-3 and e_vp < 0.05


def test_patched_reference_elastic_fwi_matches_unpatched_cpu_run(patched, golden_dir):
    import make_golden_fwi_elastic as M
    g = np.load(f"{golden_dir}/fwi_elastic_2iter.npz")
    fwi, model, prop, inp = M.build("cuda:0")
    for c in ("vx", "vz"):
        assert np.array_equal(inp["obs_" + c], g["obs_" + c]), c
    fwi.forward(iteration=2, fd_order=4, batch_size=2, checkpoint_segments=1)
    loss = np.array(fwi.iter_loss, dtype=np.float64)
    e_loss = np.abs(loss - g["iter_loss"]) / g["iter_loss"]
    e_model = {k: float(np.abs(getattr(model, k).detach().cpu().numpy() - g["final_" + k]).max()) for k in ("vp", "vs", "rho")}
    print(f"patched reference ElasticFWI on cuda vs unpatched CPU: loss {e_loss}, final model max diff {e_model}")
    assert e_loss.max() < 1e-4 and max(e_model.values()) < 0.05


def test_patched_reference_propagator_records_and_gradient(patched, golden_dir):
    """Reference AcousticModel + AcousticPropagator(device='cuda') through the patch at the acoustic example's scale (nt 1600)
    against the unpatched CPU run of tests/golden/make_golden_scale.py."""
    from ADFWI.model import AcousticModel
    from ADFWI.survey import Source, Receiver, Survey
    from ADFWI.propagator import AcousticPropagator
    from ADFWI.fwi.misfit import Misfit_waveform_L2
    g = np.load(f"{golden_dir}/acoustic_c1_scale.npz")
    nz, nx, nabc, nt, dx, dt, f0 = int(g["nz"]), int(g["nx"]), int(g["nabc"]), int(g["nt"]), float(g["dx"]), float(g["dt"]), float(g["f0"])
    s = Source(nt=nt, dt=dt, f0=f0)
    s.add_sources(src_x=g["src_x"], src_z=g["src_z"], src_wavelet=g["wavelet"], src_type="mt", src_mt=np.eye(3))
    r = Receiver(nt=nt, dt=dt)
    r.add_receivers(rcv_x=g["rcv_x"], rcv_z=g["rcv_z"], rcv_type="pr")
    m = AcousticModel(0, 0, nx, nz, dx, dx, g["vp_init"].copy(), g["rho_init"].copy(), vp_bound=None, vp_grad=True, free_surface=True,
                      abc_type="PML", nabc=nabc, device="cuda:0")
    prop = AcousticPropagator(m, Survey(source=s, receiver=r), device="cuda:0")
    prop.damp = torch.tensor(g["damp"], device="cuda:0")
    rec = prop.forward(checkpoint_segments=4)
    loss = Misfit_waveform_L2(dt=dt).forward(torch.tensor(g["obs_p"], device="cuda:0"), rec["p"])
    loss.backward()
    assert np.array_equal(rec["p"].detach().cpu().numpy(), g["rec_p"])
    e_g = rel_l2(m.vp.grad.cpu().numpy(), g["g_vp"])
    print(f"patched reference propagator at example scale: records bit-identical, g_vp rel-L2 {e_g:.2e}")
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * float(g["loss"]) and e_g <= 1e-4
