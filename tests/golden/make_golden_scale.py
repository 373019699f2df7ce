"""Example-scale golden fixtures from the UNMODIFIED reference, at the reference examples' own grid sizes AND
numbers of time steps (the small fixtures of make_golden.py stop at nt <= 260):

  acoustic_c1_scale.npz  iso-acoustic Marmousi2 example geometry (examples/acoustic/01-model-test/01-Marmousi2:
                         88 x 200 cells, dx 40 m, nabc 30 -> 148 x 260 padded, nt 1600, dt 3 ms, f0 5 Hz, free surface),
                         4 of its shots, through AcousticModel + AcousticPropagator.forward() + Misfit_waveform_L2 +
                         loss.backward(): records, loss, vp gradient.
  vti_c4_scale.npz       VTI example (examples/elastic/VTI-elastic-Anomaly/02_inversion.py:30-121: 80 x 180, dx 10 m,
                         nabc 50 -> 132 x 280 padded, nt 1000, dt 1 ms, f0 30 Hz, sources at z = 70, receivers at z = 10),
                         2 of its shots, AnisotropicElasticModel + ElasticPropagator.forward(fd_order=4) + L2 misfit on
                         vx, vz: records, loss, eps / delta / vp / vs / rho gradients.

    python tests/golden/make_golden_scale.py [acoustic] [vti]      # needs /root/reference (or $ADFWI_REF); ~5 min on 8 cores

Receivers are decimated (every 4th / 3rd) to keep the committed files small; everything else is the example's.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from adfwi_b200 import synthetic as syn  # noqa: E402  (numpy-only helpers: model shapes, wavelet)

torch.set_num_threads(os.cpu_count() or 8)


def acoustic_c1():
    ref_loader.load()
    from ADFWI.model import AcousticModel
    from ADFWI.survey import Source, Receiver, Survey
    from ADFWI.propagator import AcousticPropagator
    from ADFWI.fwi.misfit import Misfit_waveform_L2
    nz, nx, nabc, nt = 88, 200, 30, 1600
    dx = dz = 40.0
    dt, f0 = 3e-3, 5.0
    vp_true = syn.marmousi_like_vp(nz, nx)
    vp_init = syn.smooth2d(vp_true, 6)
    wav = syn.integrated_ricker(nt, dt, f0).astype(np.float32)
    src_x = np.array([2, 67, 132, 197]); src_z = np.ones(4, dtype=int)
    rcv_x = np.arange(0, nx, 4); rcv_z = np.ones(len(rcv_x), dtype=int)

    def survey():
        s = Source(nt=nt, dt=dt, f0=f0)
        s.add_sources(src_x=src_x, src_z=src_z, src_wavelet=wav, src_type="mt", src_mt=np.eye(3))
        r = Receiver(nt=nt, dt=dt)
        r.add_receivers(rcv_x=rcv_x, rcv_z=rcv_z, rcv_type="pr")
        return Survey(source=s, receiver=r)

    def model(vp, grad):
        return AcousticModel(0, 0, nx, nz, dx, dz, vp.copy(), syn.gardner_rho(vp), vp_bound=None, vp_grad=grad,
                             free_surface=True, abc_type="PML", nabc=nabc, device="cpu")

    sv = survey()
    t0 = time.time()
    true_prop = AcousticPropagator(model(vp_true, False), sv, device="cpu")
    with torch.no_grad():
        obs = true_prop.forward()["p"].detach().clone()
    m = model(vp_init, True)
    prop = AcousticPropagator(m, sv, device="cpu")
    prop.damp = true_prop.damp                       # one absorbing profile for both (vmax of the true model)
    rec = prop.forward(checkpoint_segments=4)
    loss = Misfit_waveform_L2(dt=dt).forward(obs, rec["p"])
    loss.backward()
    print(f"acoustic_c1_scale: reference forward + backward {time.time() - t0:.1f} s, loss {float(loss):.6e}")
    np.savez_compressed(
        os.path.join(HERE, "acoustic_c1_scale.npz"),
        nz=nz, nx=nx, nabc=nabc, nt=nt, dx=dx, dz=dz, dt=dt, f0=f0, vp_true=vp_true, vp_init=vp_init,
        rho_init=m.rho.detach().numpy(), wavelet=wav, src_x=src_x, src_z=src_z, rcv_x=rcv_x, rcv_z=rcv_z,
        damp=true_prop.damp.numpy(), obs_p=obs.numpy(), rec_p=rec["p"].detach().numpy(),
        rec_u=rec["u"].detach().numpy()[:, :, ::5], rec_w=rec["w"].detach().numpy()[:, :, ::5],
        illum_p=rec["forward_wavefield_p"].numpy(), loss=float(loss), g_vp=m.vp.grad.numpy())


def vti_c4():
    ref_loader.load()
    from ADFWI.model import AnisotropicElasticModel
    from ADFWI.survey import Source, Receiver, Survey
    from ADFWI.propagator import ElasticPropagator
    from ADFWI.fwi.misfit import Misfit_waveform_L2
    nz, nx, nabc, nt = 80, 180, 50, 1000
    dx = dz = 10.0
    dt, f0 = 1e-3, 30.0
    one = np.ones((nz, nx), np.float32)
    vp, vs, rho = 3000 * one, 1500 * one, 2450 * one
    eps_true, delta, gamma = 0.1 * one, -0.1 * one, 0 * one
    # the three epsilon anomalies of the example (02_inversion.py:52-77)
    zz, xx = np.mgrid[0:nz, 0:nx]
    eps_true = eps_true.copy()
    eps_true[np.sqrt((zz - nz // 2) ** 2 + (xx - nx // 2) ** 2) < 10] = 0.15
    eps_true[nz // 2 - 10:nz // 2 + 10, nx // 4 - 10:nx // 4 + 10] = 0.28
    eps_true[nz // 2 - 10:nz // 2 + 10, 3 * nx // 4 - 10:3 * nx // 4 + 15] = 0.2
    eps_init = 0.1 * one
    wav = syn.integrated_ricker(nt, dt, f0).astype(np.float32)
    src_x = np.array([41, 131]); src_z = np.array([70, 70])
    rcv_x = np.arange(0, nx, 3); rcv_z = np.full(len(rcv_x), 10)

    def survey():
        s = Source(nt=nt, dt=dt, f0=f0)
        s.add_sources(src_x=src_x, src_z=src_z, src_wavelet=wav, src_type="mt", src_mt=np.eye(3))
        r = Receiver(nt=nt, dt=dt)
        r.add_receivers(rcv_x=rcv_x, rcv_z=rcv_z, rcv_type="pr")
        return Survey(source=s, receiver=r)

    def model(eps, grad):
        return AnisotropicElasticModel(0, 0, nx, nz, dx, dz, vp=vp.copy(), vs=vs.copy(), rho=rho.copy(), eps=eps.copy(), gamma=gamma.copy(),
                                       delta=delta.copy(), vp_grad=grad, vs_grad=grad, rho_grad=grad, eps_grad=grad, gamma_grad=False,
                                       delta_grad=grad, free_surface=True, anisotropic_type="vti", abc_type="PML", abc_jerjan_alpha=0.007,
                                       auto_update_rho=False, auto_update_vp=False, nabc=nabc, device="cpu", dtype=torch.float32)

    sv = survey()
    t0 = time.time()
    true_prop = ElasticPropagator(model(eps_true, False), sv, device="cpu")
    with torch.no_grad():
        o = true_prop.forward(fd_order=4, checkpoint_segments=1)
        obs = {c: o[c].detach().clone() for c in ("vx", "vz")}
    m = model(eps_init, True)
    prop = ElasticPropagator(m, sv, device="cpu")
    rec = prop.forward(fd_order=4, checkpoint_segments=4)
    fn = Misfit_waveform_L2(dt=dt)
    loss = fn.forward(obs["vx"], rec["vx"]) + fn.forward(obs["vz"], rec["vz"])
    loss.backward()
    print(f"vti_c4_scale: reference forward + backward {time.time() - t0:.1f} s, loss {float(loss):.6e}")
    np.savez_compressed(
        os.path.join(HERE, "vti_c4_scale.npz"),
        nz=nz, nx=nx, nabc=nabc, nt=nt, dx=dx, dz=dz, dt=dt, f0=f0, vp=vp, vs=vs, rho=rho, eps_true=eps_true, eps_init=eps_init,
        delta=delta, wavelet=wav, src_x=src_x, src_z=src_z, rcv_x=rcv_x, rcv_z=rcv_z, bcx=prop.bcx.numpy(), bcz=prop.bcz.numpy(),
        obs_vx=obs["vx"].numpy(), obs_vz=obs["vz"].numpy(), rec_vx=rec["vx"].detach().numpy(), rec_vz=rec["vz"].detach().numpy(),
        rec_txx=rec["txx"].detach().numpy()[:, :, ::6], rec_tzz=rec["tzz"].detach().numpy()[:, :, ::6], rec_txz=rec["txz"].detach().numpy()[:, :, ::6],
        loss=float(loss), **{f"g_{k}": getattr(m, k).grad.numpy() for k in ("vp", "vs", "rho", "eps", "delta")})


if __name__ == "__main__":
    which = sys.argv[1:] or ["acoustic", "vti"]
    if "acoustic" in which:
        acoustic_c1()
    if "vti" in which:
        vti_c4()
