"""Synthetic models / surveys of SURVEY.md section 8(d), as light containers exposing exactly the
attributes the propagators read from the upstream ``AbstractModel`` / ``Survey`` objects (those
classes are outside the hot path and are not rebuilt here; the real ones work unchanged)."""
import numpy as np
import torch


def ricker(nt, dt, f0, t0=None):
    """Ricker wavelet with delay t0 (default 1.2/f0), as ADFWI/utils/wavelets.py:4-15 evaluates it."""
    t = np.arange(nt) * dt
    t0 = 1.2 / f0 if t0 is None else t0
    a = (np.pi * f0 * (t - t0)) ** 2
    return (1.0 - 2.0 * a) * np.exp(-a)


def integrated_ricker(nt, dt, f0):
    """Time-integrated Ricker (cumulative trapezoid with a leading zero), the source the upstream
    examples feed to the stress-velocity propagators."""
    w = ricker(nt, dt, f0)
    out = np.zeros(nt)
    out[1:] = np.cumsum(0.5 * (w[1:] + w[:-1])) * dt
    return out


def marmousi_like_vp(nz, nx):
    """vp(z,x) = 1500 + 3000 z/zmax + 300 sin(2 pi 3 x/xmax + 4 z/zmax), clipped to [1500,4700] m/s."""
    z = (np.arange(nz) / max(nz - 1, 1))[:, None]
    x = (np.arange(nx) / max(nx - 1, 1))[None, :]
    vp = 1500.0 + 3000.0 * z + 300.0 * np.sin(2 * np.pi * 3 * x + 4 * z)
    return np.clip(vp, 1500.0, 4700.0).astype(np.float32)


def gardner_rho(vp):
    return (310.0 * np.asarray(vp, dtype=np.float64) ** 0.25).astype(np.float32)   # acoustic_model.py:115


def smooth2d(a, sigma):
    from scipy.ndimage import gaussian_filter
    return gaussian_filter(np.asarray(a, dtype=np.float64), sigma, mode="nearest").astype(np.float32)


class Source:
    def __init__(self, loc, wavelet, nt, dt, f0, moment_tensor=None):
        self.loc = np.asarray(loc, dtype=np.int64).reshape(-1, 2)      # (x, z) grid indices
        self.num = len(self.loc)
        self.nt, self.dt, self.f0 = nt, dt, f0
        w = np.asarray(wavelet, dtype=np.float32)
        self.wavelet = np.broadcast_to(w, (self.num, nt)).copy() if w.ndim == 1 else w
        if moment_tensor is None:
            moment_tensor = np.broadcast_to(np.eye(3, dtype=np.float32), (self.num, 3, 3)).copy()
        self.moment_tensor = np.asarray(moment_tensor, dtype=np.float32)

    def get_loc(self): return self.loc
    def get_wavelet(self): return self.wavelet
    def get_moment_tensor(self): return self.moment_tensor


class Receiver:
    def __init__(self, loc):
        self.loc = np.asarray(loc, dtype=np.int64).reshape(-1, 2)
        self.num = len(self.loc)

    def get_loc(self): return self.loc


class Survey:
    def __init__(self, source, receiver):
        self.source, self.receiver = source, receiver


class AcousticGridModel(torch.nn.Module):
    """vp / rho on an (nz,nx) grid with the attributes AcousticPropagator reads
    (ADFWI/model/acoustic_model.py).  ``auto_update_rho``: rho <- 310 vp^0.25 detached on every
    forward(), as the upstream model does (acoustic_model.py:111-130)."""

    def __init__(self, vp, rho=None, dx=10.0, dz=10.0, ox=0.0, oz=0.0, nabc=50, free_surface=True,
                 abc_type="PML", vp_grad=True, rho_grad=False, auto_update_rho=True, device="cuda"):
        super().__init__()
        vp = torch.as_tensor(np.asarray(vp), dtype=torch.float32)
        self.nz, self.nx = vp.shape
        self.dx, self.dz, self.ox, self.oz = dx, dz, ox, oz
        self.nabc, self.free_surface, self.abc_type = nabc, free_surface, abc_type
        self.abc_jerjan_alpha = 0.0053
        self.auto_update_rho = auto_update_rho
        self.vp = torch.nn.Parameter(vp.to(device), requires_grad=vp_grad)
        rho = gardner_rho(vp.numpy()) if rho is None else np.asarray(rho)
        self.rho = torch.nn.Parameter(torch.as_tensor(rho, dtype=torch.float32).to(device), requires_grad=rho_grad)

    def forward(self):
        if self.auto_update_rho:
            with torch.no_grad():
                self.rho.copy_(310.0 * self.vp.detach() ** 0.25)
        return None


def surface_survey(nx, ns, nr, nt, dt, f0, src_z=1, rcv_z=1):
    """ns sources evenly spaced along x at depth src_z; nr receivers evenly spaced at depth rcv_z."""
    sx = np.round(np.linspace(2, nx - 3, ns)).astype(np.int64) if ns > 1 else np.array([nx // 2])
    rx = np.round(np.linspace(0, nx - 1, nr)).astype(np.int64)
    src = Source(np.stack([sx, np.full(ns, src_z)], 1), integrated_ricker(nt, dt, f0), nt, dt, f0)
    rcv = Receiver(np.stack([rx, np.full(nr, rcv_z)], 1))
    return Survey(src, rcv)


class ElasticGridModel(torch.nn.Module):
    """vp, vs, rho (+ Thomsen eps, delta, gamma) on an (nz,nx) grid exposing what ElasticPropagator
    reads after ``forward()``: ``lamu, lam, bx, bz, CC`` (21 moduli; C11,C13,C33,C55 used).

    The parameterisation restates the physics of ADFWI/model/parameters.py in plain torch so that
    autograd carries the propagator's coefficient-plane gradients back to the model parameters:
    C33 = vp^2 rho, C44 = vs^2 rho, C11 = C33 (1+2 eps), C66 = C44 (1+2 gamma),
    C13 = sqrt(2 C33 (C33-C44) delta + (C33-C44)^2) - C44 (:102-107), C55 = C44 (VTI / isotropic),
    b = 1/rho, bx/bz = two-point averages of b, C55 <- 0.2*(five-point average counting the lower
    neighbour twice) (:199-212).  Isotropic models are the eps = delta = gamma = 0 case."""

    def __init__(self, vp, vs, rho, eps=None, delta=None, gamma=None, dx=10.0, dz=10.0, ox=0.0, oz=0.0,
                 nabc=50, free_surface=True, abc_type="PML", requires_grad=("vp", "vs", "rho"), device="cuda"):
        super().__init__()
        t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).to(device)
        vp = t(vp)
        self.nz, self.nx = vp.shape
        self.dx, self.dz, self.ox, self.oz = dx, dz, ox, oz
        self.nabc, self.free_surface, self.abc_type = nabc, free_surface, abc_type
        self.abc_jerjan_alpha = 0.0053
        zeros = torch.zeros_like(vp)
        vals = dict(vp=vp, vs=t(vs), rho=t(rho), eps=zeros.clone() if eps is None else t(eps),
                    delta=zeros.clone() if delta is None else t(delta), gamma=zeros.clone() if gamma is None else t(gamma))
        for k, v in vals.items():
            setattr(self, k, torch.nn.Parameter(v, requires_grad=k in requires_grad))
        self.lamu = self.lam = self.bx = self.bz = self.CC = None

    def forward(self):
        vp, vs, rho = self.vp, self.vs, self.rho
        C33 = vp ** 2 * rho
        C44 = vs ** 2 * rho
        C11 = C33 * (1 + 2 * self.eps)
        C66 = C44 * (1 + 2 * self.gamma)
        C13 = torch.sqrt(2 * C33 * (C33 - C44) * self.delta + (C33 - C44) ** 2) - C44
        self.lamu, self.lam = C33, C33 - 2 * C44
        b = 1 / rho
        nz, nx = self.nz, self.nx
        self.bx = 0.5 * (b[:, 0:nx - 1] + b[:, 1:nx])
        self.bz = 0.5 * (b[0:nz - 1, :] + b[1:nz, :])
        C55 = C44
        C55s = 0.2 * (C55[1:nz - 1, 1:nx - 1] + C55[2:nz, 1:nx - 1] + C55[1:nz - 1, 2:nx] + C55[2:nz, 1:nx - 1] + C55[2:nz, 2:nx])
        zero = torch.zeros_like(vp)
        CC = [zero] * 21
        CC[0], CC[2], CC[11], CC[18] = C11, C13, C33, C55s
        CC[15], CC[20] = C44, C66
        self.CC = CC
        return None
