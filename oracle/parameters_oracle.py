"""numpy restatement of the reference's elastic parameterisation (TEST INFRASTRUCTURE ONLY):
thomsen_to_elastic_moduli (ADFWI/model/parameters.py:71-107), C55 = C44 (:150,171), b = 1/rho (:66), parameter_staggered_grid
(:184-213), in float32 with the eager association, and the transpose of that chain in float64.  Pinned by
tests/test_parameters_oracle.py to the `in_*` planes and the model-level gradients of tests/golden/elastic_*.npz (outputs of the
unmodified reference's own chain and autograd)."""
import numpy as np

f32 = np.float32


def planes(vp, vs, rho, eps, delta):
    vp, vs, rho, eps, delta = (np.asarray(a, f32) for a in (vp, vs, rho, eps, delta))
    nz, nx = vp.shape
    C33 = (vp * vp) * rho
    C44 = (vs * vs) * rho
    C11 = C33 * (f32(1) + f32(2) * eps)
    A = C33 - C44
    C13 = np.sqrt(((f32(2) * C33) * A) * delta + A * A) - C44
    b = f32(1) / rho
    bx = f32(0.5) * (b[:, 0:nx - 1] + b[:, 1:nx])
    bz = f32(0.5) * (b[0:nz - 1, :] + b[1:nz, :])
    C55 = f32(0.2) * ((((C44[1:nz - 1, 1:nx - 1] + C44[2:nz, 1:nx - 1]) + C44[1:nz - 1, 2:nx]) + C44[2:nz, 1:nx - 1]) + C44[2:nz, 2:nx])
    return dict(C11=C11, C13=C13, C33=C33, C55=C55, bx=bx, bz=bz)


def planes_T(vp, vs, rho, eps, delta, g):
    """Transpose of planes(): cotangents g[name] (own shapes) -> (g_vp, g_vs, g_rho, g_eps, g_delta), float64."""
    vp, vs, rho, eps, delta = (np.asarray(a, np.float64) for a in (vp, vs, rho, eps, delta))
    g = {k: np.asarray(v, np.float64) for k, v in g.items()}
    nz, nx = vp.shape
    C33, C44 = vp * vp * rho, vs * vs * rho
    A = C33 - C44
    R = np.sqrt(2 * C33 * A * delta + A * A)
    gb = np.zeros((nz, nx)); g44 = np.zeros((nz, nx))
    gb[:, 0:nx - 1] += 0.5 * g["bx"]; gb[:, 1:nx] += 0.5 * g["bx"]
    gb[0:nz - 1, :] += 0.5 * g["bz"]; gb[1:nz, :] += 0.5 * g["bz"]
    g55 = 0.2 * g["C55"]
    g44[1:nz - 1, 1:nx - 1] += g55; g44[2:nz, 1:nx - 1] += 2 * g55; g44[1:nz - 1, 2:nx] += g55; g44[2:nz, 2:nx] += g55
    t33 = g["C33"] + g["C11"] * (1 + 2 * eps) + g["C13"] * (delta * (A + C33) + A) / R
    t44 = g44 + g["C13"] * (-(C33 * delta + A) / R - 1)
    return (t33 * 2 * vp * rho, t44 * 2 * vs * rho, t33 * vp * vp + t44 * vs * vs - gb / rho ** 2, g["C11"] * 2 * C33, g["C13"] * C33 * A / R)
