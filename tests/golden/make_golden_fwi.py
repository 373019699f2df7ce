"""Iteration-level golden fixture: three iterations of the UNMODIFIED reference's AcousticFWI loop on CPU
(ADFWI/fwi/acoustic_fwi.py:115-200: shot batches, per-trace max normalisation, Misfit_waveform_L2, backward,
GradProcessor with the "Marine" mute + illumination preconditioner + smoothing + max-normalisation, SGD step, StepLR).

    python tests/golden/make_golden_fwi.py        # needs /root/reference (or $ADFWI_REF)

Stores the inputs (true / initial model, geometry, wavelet, observed records) and, per iteration, the loss, the
processed gradient and the updated vp, for tests/test_fwi_iterations_gpu.py.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_loader  # noqa: E402

torch.set_num_threads(8)


def ricker_integral(nt, dt, f0):
    t = np.arange(nt) * dt
    a = (np.pi * f0 * (t - 1.2 / f0)) ** 2
    return np.cumsum((1 - 2 * a) * np.exp(-a)) * dt


def build(device="cpu"):
    """The reference objects of the fixture on ``device``: returns (AcousticFWI instance, model, propagator, inputs dict).
    Shared with tests/test_reference_patch_gpu.py, which runs the same script on the GPU after adfwi_b200.patch()."""
    ref_loader.load()
    from ADFWI.model import AcousticModel
    from ADFWI.survey import Source, Receiver, Survey, SeismicData
    from ADFWI.propagator import AcousticPropagator, GradProcessor
    from ADFWI.fwi import AcousticFWI
    from ADFWI.fwi.misfit import Misfit_waveform_L2
    import ADFWI.fwi.acoustic_fwi as _afwi
    _afwi.NLCG = type("NLCG", (), {})          # `ncg_optimizer` is stubbed (not installed): isinstance() needs a real type

    nz, nx, nabc, nt = 40, 64, 12, 260
    dx = dz = 10.0
    dt = 1e-3
    f0 = 22.0
    z = np.linspace(0, 1, nz)[:, None]; x = np.linspace(0, 1, nx)[None, :]
    vp_true = (1500 + 1500 * z + 120 * np.sin(2 * np.pi * 1.5 * x + 2 * z)).astype(np.float32)
    vp_true[:6] = 1500.0                                   # water layer
    vp_true[18:24, 24:40] += 300.0                         # anomaly
    from scipy.ndimage import gaussian_filter
    vp_init = gaussian_filter(vp_true.astype(np.float64), 4, mode="nearest").astype(np.float32)
    vp_init[:6] = 1500.0
    rho = (310.0 * vp_init.astype(np.float64) ** 0.25).astype(np.float32)
    wav = ricker_integral(nt, dt, f0).astype(np.float32)
    src_x = np.array([6, 19, 32, 45, 58]); src_z = np.full(5, 1)
    rcv_x = np.arange(0, nx, 2); rcv_z = np.full(len(rcv_x), 1)

    def survey():
        s = Source(nt=nt, dt=dt, f0=f0)
        s.add_sources(src_x=src_x, src_z=src_z, src_wavelet=wav, src_type="mt", src_mt=np.eye(3))
        r = Receiver(nt=nt, dt=dt)
        r.add_receivers(rcv_x=rcv_x, rcv_z=rcv_z, rcv_type="pr")
        return Survey(source=s, receiver=r)

    def model(vp, grad):
        return AcousticModel(0, 0, nx, nz, dx, dz, vp.copy(), (310.0 * vp.astype(np.float64) ** 0.25).astype(np.float32),
                             vp_bound=None, vp_grad=grad, free_surface=True, abc_type="PML", nabc=nabc, device=device)

    sv = survey()
    true_prop = AcousticPropagator(model(vp_true, False), sv, device=device)
    with torch.no_grad():
        obs = true_prop.forward()
    data = SeismicData(sv)
    data.record_data({"p": obs["p"], "u": obs["u"], "w": obs["w"]})
    obs_p = np.array(data.data["p"], dtype=np.float32)

    m = model(vp_init, True)
    prop = AcousticPropagator(m, sv, device=device)
    opt = torch.optim.SGD(m.parameters(), lr=0.01)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.5)
    gp = GradProcessor(grad_mute=6, grad_smooth=2, grad_mask=None, norm_grad=True, forw_illumination=True, marine_or_land="Marine")
    fwi = AcousticFWI(propagator=prop, model=m, optimizer=opt, scheduler=sched, loss_fn=Misfit_waveform_L2(dt=1), obs_data=data,
                      gradient_processor=gp, waveform_normalize=True, cache_result=True, save_fig_epoch=-1)
    inputs = dict(nz=nz, nx=nx, nabc=nabc, nt=nt, dx=dx, dz=dz, dt=dt, f0=f0, vp_true=vp_true, vp_init=vp_init, rho=rho, wavelet=wav,
                  src_x=src_x, src_z=src_z, rcv_x=rcv_x, rcv_z=rcv_z, obs_p=obs_p)
    return fwi, m, prop, inputs


def main():
    fwi, m, prop, inp = build("cpu")
    nz, nx, nabc, nt, dx, dz, dt, f0 = (inp[k] for k in ("nz", "nx", "nabc", "nt", "dx", "dz", "dt", "f0"))
    vp_true, vp_init, rho, wav, src_x, src_z, rcv_x, rcv_z, obs_p = (inp[k] for k in ("vp_true", "vp_init", "rho", "wavelet", "src_x", "src_z", "rcv_x", "rcv_z", "obs_p"))
    fwi.forward(iteration=3, batch_size=2, checkpoint_segments=1)
    out = dict(nz=nz, nx=nx, nabc=nabc, nt=nt, dx=dx, dz=dz, dt=dt, f0=f0, vp_true=vp_true, vp_init=vp_init, rho=rho, wavelet=wav,
               src_x=src_x, src_z=src_z, rcv_x=rcv_x, rcv_z=rcv_z, obs_p=obs_p, damp=np.array(prop.damp.cpu().numpy()),
               iter_vp=np.stack(fwi.iter_vp), iter_grad=np.stack(fwi.iter_vp_grad), iter_loss=np.array(fwi.iter_loss, dtype=np.float64),
               lr=0.01, step_size=2, gamma=0.5, batch_size=2, grad_mute=6, grad_smooth=2)
    np.savez_compressed(os.path.join(HERE, "fwi_acoustic_3iter.npz"), **out)
    print("loss", out["iter_loss"], "max |dvp|", [float(np.abs(v - vp_init).max()) for v in out["iter_vp"]])


if __name__ == "__main__":
    main()
