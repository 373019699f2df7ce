"""Fused record post-processing (normalisation + misfit + adjoint source) and the regularisers on the device, through the C ABI
(adfwi_misfit_*, adfwi_regularization_*), against fixtures of the UNMODIFIED reference classes and against the numpy oracle on
larger, ragged cases."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _fused(syn, obs, kind, norm, dt, scale=1.0):
    from adfwi_b200.fwi import misfit as M
    cls = M.Misfit_waveform_L2 if kind == 0 else M.Misfit_global_correlation
    s = torch.tensor(syn, device="cuda:0", requires_grad=True)
    loss = cls(dt=dt, normalize=bool(norm)).forward(torch.tensor(obs, device="cuda:0"), s)
    (loss * scale).backward()
    return float(loss), s.grad.cpu().numpy()


@pytest.mark.parametrize("kind", ["l2", "gc"])
@pytest.mark.parametrize("norm", [0, 1])
@pytest.mark.parametrize("dts", ["1", "s"])
def test_fused_misfit_matches_reference(golden_dir, kind, norm, dts):
    g = np.load(f"{golden_dir}/objective_misfit.npz")
    dt = 1.0 if dts == "1" else float(g["dt_s"])
    loss, grad = _fused(g["syn"], g["obs"], 0 if kind == "l2" else 1, norm, dt)
    tag = f"{kind}_n{norm}_dt{dts}"
    assert abs(loss - float(g["loss_" + tag])) <= 2e-6 * abs(float(g["loss_" + tag])), tag
    assert rel_l2(grad, g["g_" + tag]) <= 2e-5, tag


@pytest.mark.parametrize("shape", [(2, 1000, 130), (5, 333, 1700), (1, 64, 3)])
def test_fused_misfit_matches_oracle_on_larger_records(shape):
    """Several time chunks per trace, ragged receiver counts, an upstream gradient different from 1."""
    from oracle import objective_oracle as OO
    rng = np.random.default_rng(3)
    ns, nt, nr = shape
    t = np.arange(nt)[None, :, None]
    syn = (np.sin(0.05 * t * (1 + rng.random((ns, 1, nr)))) * np.exp(-((t - nt / 2) / (nt / 4)) ** 2) * 1e-4 * (1 + rng.random((ns, 1, nr))) +
           1e-7 * rng.standard_normal((ns, nt, nr))).astype(np.float32)
    obs = (np.sin(0.05 * t * (1 + rng.random((ns, 1, nr))) + 0.3) * np.exp(-((t - nt / 2) / (nt / 4)) ** 2) + 1e-3 * rng.standard_normal((ns, nt, nr))).astype(np.float32)
    obs = (obs / np.abs(obs).max(axis=1, keepdims=True)).astype(np.float32)
    for kind in (0, 1):
        for norm in (0, 1):
            loss, grad = _fused(syn, obs, kind, norm, 1.0, scale=2.5)
            rl, rg = OO.misfit(syn, obs, kind, bool(norm), 1.0)
            assert abs(loss - rl) <= 2e-6 * abs(rl), (kind, norm)
            assert rel_l2(grad, 2.5 * rg) <= 2e-5, (kind, norm)


def test_fused_misfit_drives_the_propagator(golden_dir):
    """loss.backward() through the fused misfit hands the adjoint source to the acoustic propagator: same vp gradient as the
    eager torch normalisation + L2 expression of the reference driver."""
    from adfwi_b200 import fwi, synthetic as syn
    from adfwi_b200.fwi import misfit as M
    from adfwi_b200.propagator import AcousticPropagator
    g = np.load(f"{golden_dir}/fwi_acoustic_3iter.npz")
    dev = torch.device("cuda:0")
    nt, dt, f0 = int(g["nt"]), float(g["dt"]), float(g["f0"])
    obs = torch.tensor(g["obs_p"], device=dev)
    obs = obs / torch.max(torch.abs(obs), dim=1, keepdim=True).values
    grads = []
    for fused in (False, True):
        model = syn.AcousticGridModel(g["vp_init"], dx=float(g["dx"]), dz=float(g["dz"]), nabc=int(g["nabc"]), free_surface=True, vp_grad=True, device=dev)
        prop = AcousticPropagator(model, syn.Survey(syn.Source(np.stack([g["src_x"], g["src_z"]], 1), g["wavelet"], nt, dt, f0),
                                                    syn.Receiver(np.stack([g["rcv_x"], g["rcv_z"]], 1))), device=dev)
        prop.damp = torch.tensor(g["damp"], device=dev)
        rec = prop.forward()["p"]
        if fused:
            loss = M.normalized_misfit(rec, obs, M.Misfit_waveform_L2(dt=1))
        else:
            loss = fwi.l2_waveform_misfit(obs, rec / torch.max(torch.abs(rec), dim=1, keepdim=True).values, 1.0)
        loss.backward()
        grads.append((float(loss), model.vp.grad.cpu().numpy()))
    assert abs(grads[0][0] - grads[1][0]) <= 2e-6 * abs(grads[0][0])
    assert rel_l2(grads[1][1], grads[0][1]) <= 2e-5


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_regularizers_match_reference(golden_dir, kind):
    from adfwi_b200.fwi import regularization as R
    g = np.load(f"{golden_dir}/objective_regularization.npz")
    nz, nx = g["m"].shape
    cls = (R.TV_1order, R.Tikhonov_1order, R.TV_2order, R.Tikhonov_2order)[kind]
    reg = cls(nx, nz, float(g["dx"]), float(g["dz"]), float(g["alphax"]), float(g["alphaz"]), step_size=1000, gamma=1)
    m = torch.tensor(g["m"], device="cuda:0", requires_grad=True)
    v = reg.forward(m)
    (3.0 * v).backward()
    assert abs(float(v) - float(g[f"value_{kind}"])) <= 2e-6 * float(g[f"value_{kind}"])
    assert rel_l2(m.grad.cpu().numpy(), 3.0 * g[f"g_{kind}"]) <= 2e-5
    assert reg.iter == (0 if kind == 2 else 1)          # TV_2order does not advance its step counter upstream


def test_regularizers_match_oracle_on_a_model_sized_plane():
    from adfwi_b200.fwi import regularization as R
    from oracle import objective_oracle as OO
    rng = np.random.default_rng(5)
    nz, nx = 350, 1700
    m = (2000 + 1500 * np.linspace(0, 1, nz)[:, None] + 100 * rng.standard_normal((nz, nx))).astype(np.float32)
    for kind, cls in enumerate((R.TV_1order, R.Tikhonov_1order, R.TV_2order, R.Tikhonov_2order)):
        reg = cls(nx, nz, 10.0, 10.0, 1e-4, 2e-4, step_size=1, gamma=0.5)
        reg.iter = 2                                      # decayed factors: alpha * 0.5^2
        mt = torch.tensor(m, device="cuda:0", requires_grad=True)
        v = reg.forward(mt)
        v.backward()
        rv, rg = OO.regularization(m, kind, 10.0, 10.0, 1e-4 * 0.25, 2e-4 * 0.25)
        assert abs(float(v) - rv) <= 5e-6 * rv, kind
        assert rel_l2(mt.grad.cpu().numpy(), rg) <= 2e-5, kind
