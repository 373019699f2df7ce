"""Absorbing-boundary profiles, built once per propagator on the host (numpy float64).

Same names, arguments and return layouts as ``ADFWI/propagator/boundary_condition.py``:
``bc_pml`` (:15-47), ``bc_sincos`` (:50-68), ``bc_gerjan`` (:71-95), ``bc_pml_xz`` (:98-120).
Bit-for-bit agreement with the reference is pinned by tests/golden/boundary_profiles.npz.
All return arrays are laid out (nz_pml, nx_pml).
"""
from math import log

import numpy as np

_R = 1e-6   # target reflection coefficient used by both PML flavours


def _padded_shape(nx, nz, pml, free_surface):
    return nz + (pml if free_surface else 2 * pml), nx + 2 * pml


def bc_pml(nx, nz, dx, dz, pml, vmax, free_surface=True):
    """Acoustic damping plane: quadratic ramp kappa*(i*dx/a)^2 (a = (pml-1)*dx) towards the left,
    right and bottom (and top unless ``free_surface``) edges; top/bottom ramps cover the trapezoid
    x in [pml-iz, nx_pml-pml+iz) so the side ramps win in the corners."""
    nz_pml, nx_pml = _padded_shape(nx, nz, pml, free_surface)
    a = (pml - 1) * dx
    kappa = -3.0 * vmax * log(_R) / (2.0 * a)
    xa = np.arange(pml) * dx / a
    ramp = kappa * xa * xa
    out = np.zeros((nz_pml, nx_pml))
    out[:, pml - 1::-1] = ramp[None, :] if pml > 0 else 0.0          # column pml-1-i  <- ramp[i]
    out[:, nx_pml - pml:] = ramp[None, :]                            # column nx_pml-pml+i
    for iz in range(pml):
        lo, hi = pml - iz, nx_pml - pml + iz
        if lo >= hi:
            continue
        if not free_surface:
            out[pml - iz - 1, lo:hi] = ramp[iz]
        out[nz_pml - pml + iz, lo:hi] = ramp[iz]
    return out


def bc_sincos(nx, nz, dx, dz, pml, free_surface=False):
    """Multiplicative sin^2 sponge: weights sin(pi/2 * i/pml)^2 multiplied in from each edge."""
    nz_pml, nx_pml = _padded_shape(nx, nz, pml, free_surface)
    out = np.ones((nz_pml, nx_pml))
    for i in range(pml):
        wgt = np.sin(np.pi / 2 * i / pml) ** 2
        if not free_surface:
            out[i, :] *= wgt
        out[nz_pml - 1 - i, :] *= wgt
        out[:, i] *= wgt
        out[:, nx_pml - 1 - i] *= wgt
    return out


def bc_gerjan(nx, nz, dx, dz, pml, alpha=0.0053, free_surface=True):
    """Cerjan et al. (1985) sponge: nested frames with weight exp(-(alpha*(pml-k))^2), frame k
    (1 = outermost) painted left, bottom, right (and top without a free surface) in that order."""
    nz_pml, nx_pml = _padded_shape(nx, nz, pml, free_surface)
    wt = np.exp(-(alpha * (pml - np.arange(1, pml + 1))) ** 2)
    out = np.ones((nz_pml, nx_pml))
    for k in range(1, pml + 1):
        wk = wt[k - 1]
        if free_surface:
            out[:nz_pml - k + 1, k - 1] = wk
            out[nz_pml - k, k - 1:nx_pml - k + 1] = wk
            out[:nz_pml - k + 1, nx_pml - k] = wk
        else:
            out[k - 1:nz_pml - k + 1, k - 1] = wk
            out[k:nz_pml - k + 1, nx_pml - k] = wk
            out[nz_pml - k, k - 1:nx_pml - k + 1] = wk
            out[k - 1, k - 1:nx_pml - k + 1] = wk
    return out


def bc_pml_xz(nx, nz, dx, dz, pml, vmax, free_surface=True):
    """Split-PML profiles: BCx depends on x only, BCz on z only; value (pml-k+1)^2 * ppml / h at
    distance k-1 from the edge, ppml = -ln(R)*3*vmax/(2*pml^3)."""
    nz_pml, nx_pml = _padded_shape(nx, nz, pml, free_surface)
    ppml = -np.log(_R) * 3 * vmax / (2 * pml ** 3)
    k = np.arange(1, pml + 1)
    prof_x = (pml - k + 1) ** 2 * ppml / dx
    prof_z = (pml - k + 1) ** 2 * ppml / dz
    bcx = np.zeros((nz_pml, nx_pml))
    bcz = np.zeros((nz_pml, nx_pml))
    bcx[:, k - 1] = prof_x[None, :]
    bcx[:, nx_pml - k] = prof_x[None, :]
    bcz[nz_pml - k, :] = prof_z[:, None]
    if not free_surface:
        bcz[k - 1, :] = prof_z[:, None]
    return bcx, bcz
