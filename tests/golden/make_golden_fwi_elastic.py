"""Iteration-level golden fixture, elastic: two iterations of the UNMODIFIED reference's ElasticFWI loop on CPU
(ADFWI/fwi/elastic_fwi.py:196-320: shot batches, per-trace max normalisation of vx / vz, Misfit_waveform_L2, backward
through the split-PML O(2,4) propagator and the vp/vs/rho parameterisation, GradProcessor per parameter, SGD, StepLR).

    python tests/golden/make_golden_fwi_elastic.py        # needs /root/reference (or $ADFWI_REF)
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
    """The reference objects of the fixture on ``device``: returns (ElasticFWI instance, model, propagator, inputs dict).
    Shared with tests/test_reference_patch_gpu.py, which runs the same script on the GPU after adfwi_b200.patch()."""
    ref_loader.load()
    from ADFWI.model import IsotropicElasticModel
    from ADFWI.survey import Source, Receiver, Survey, SeismicData
    from ADFWI.propagator import ElasticPropagator, GradProcessor
    from ADFWI.fwi import ElasticFWI
    from ADFWI.fwi.misfit import Misfit_waveform_L2
    import ADFWI.fwi.elastic_fwi as _efwi
    _efwi.NLCG = type("NLCG", (), {})          # `ncg_optimizer` is stubbed (not installed): isinstance() needs a real type

    nz, nx, nabc, nt = 36, 56, 10, 220
    dx = dz = 10.0
    dt = 1e-3
    f0 = 20.0
    z = np.linspace(0, 1, nz)[:, None]; x = np.linspace(0, 1, nx)[None, :]
    vp_true = (2000 + 1200 * z + 100 * np.sin(2 * np.pi * 1.5 * x + 2 * z)).astype(np.float32)
    vp_true[14:20, 20:36] += 250.0
    from scipy.ndimage import gaussian_filter
    vp_init = gaussian_filter(vp_true.astype(np.float64), 4, mode="nearest").astype(np.float32)
    mk_vs = lambda v: (v / np.sqrt(3.0)).astype(np.float32)
    mk_rho = lambda v: (310.0 * v.astype(np.float64) ** 0.25).astype(np.float32)
    wav = ricker_integral(nt, dt, f0).astype(np.float32)
    src_x = np.array([8, 28, 48]); src_z = np.full(3, 2)
    rcv_x = np.arange(1, nx, 2); rcv_z = np.full(len(rcv_x), 2)

    def survey():
        s = Source(nt=nt, dt=dt, f0=f0)
        s.add_sources(src_x=src_x, src_z=src_z, src_wavelet=wav, src_type="mt", src_mt=np.eye(3))
        r = Receiver(nt=nt, dt=dt)
        r.add_receivers(rcv_x=rcv_x, rcv_z=rcv_z, rcv_type="pr")
        return Survey(source=s, receiver=r)

    def model(vp, grad):
        return IsotropicElasticModel(0, 0, nx, nz, dx, dz, vp.copy(), mk_vs(vp), mk_rho(vp), vp_grad=grad, vs_grad=grad, rho_grad=grad,
                                     free_surface=True, abc_type="PML", nabc=nabc, auto_update_rho=False, auto_update_vp=False, device=device)

    sv = survey()
    true_prop = ElasticPropagator(model(vp_true, False), sv, device=device)
    with torch.no_grad():
        obs = true_prop.forward(fd_order=4)
    data = SeismicData(sv)
    data.record_data({k: obs[k] for k in ("txx", "tzz", "txz", "vx", "vz")})
    obs_np = {k: np.array(data.data[k], dtype=np.float32) for k in ("vx", "vz")}

    m = model(vp_init, True)
    prop = ElasticPropagator(m, sv, device=device)
    opt = torch.optim.SGD(m.parameters(), lr=0.01)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.5)
    gp = GradProcessor(grad_mute=4, grad_smooth=2, grad_mask=None, norm_grad=True, forw_illumination=True, marine_or_land="Marine")
    fwi = ElasticFWI(propagator=prop, model=m, optimizer=opt, scheduler=sched, loss_fn=Misfit_waveform_L2(dt=1), obs_data=data,
                     gradient_processor=gp, waveform_normalize=True, cache_result=True, cache_gradient=True, save_fig_epoch=-1,
                     inversion_component=["vx", "vz"])
    inputs = dict(nz=nz, nx=nx, nabc=nabc, nt=nt, dx=dx, dz=dz, dt=dt, f0=f0, vp_init=vp_init, vs_init=mk_vs(vp_init), rho_init=mk_rho(vp_init),
                  wavelet=wav, src_x=src_x, src_z=src_z, rcv_x=rcv_x, rcv_z=rcv_z, obs_vx=obs_np["vx"], obs_vz=obs_np["vz"])
    return fwi, m, prop, inputs


def main():
    fwi, m, prop, inp = build("cpu")
    nz, nx, nabc, nt, dx, dz, dt, f0 = (inp[k] for k in ("nz", "nx", "nabc", "nt", "dx", "dz", "dt", "f0"))
    vp_init, wav, src_x, src_z, rcv_x, rcv_z = (inp[k] for k in ("vp_init", "wavelet", "src_x", "src_z", "rcv_x", "rcv_z"))
    mk_vs = lambda v: inp["vs_init"]; mk_rho = lambda v: inp["rho_init"]
    obs_np = {"vx": inp["obs_vx"], "vz": inp["obs_vz"]}
    fwi.forward(iteration=2, fd_order=4, batch_size=2, checkpoint_segments=1)
    final = {k: getattr(m, k).detach().cpu().numpy().copy() for k in ("vp", "vs", "rho")}
    out = dict(nz=nz, nx=nx, nabc=nabc, nt=nt, dx=dx, dz=dz, dt=dt, f0=f0, vp_init=vp_init, vs_init=mk_vs(vp_init), rho_init=mk_rho(vp_init),
               wavelet=wav, src_x=src_x, src_z=src_z, rcv_x=rcv_x, rcv_z=rcv_z, obs_vx=obs_np["vx"], obs_vz=obs_np["vz"],
               bcx=np.array(prop.bcx.cpu().numpy()), bcz=np.array(prop.bcz.cpu().numpy()),
               iter_loss=np.array(fwi.iter_loss, dtype=np.float64), final_vp=final["vp"], final_vs=final["vs"], final_rho=final["rho"],
               lr=0.01, step_size=1, gamma=0.5, batch_size=2, grad_mute=4, grad_smooth=2)
    for k in ("vp", "vs", "rho"):
        gl = getattr(fwi, "iter_" + k + "_grad", None)
        if gl is not None and len(gl):
            out["iter_grad_" + k] = np.stack(gl)
    np.savez_compressed(os.path.join(HERE, "fwi_elastic_2iter.npz"), **out)
    print("loss", out["iter_loss"], {k: float(np.abs(final[k] - out[k + "_init"]).max()) for k in ("vp", "vs", "rho")}, [k for k in out if k.startswith("iter_grad")])


if __name__ == "__main__":
    main()
