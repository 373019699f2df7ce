"""Fused elastic parameterisation and padding on the device (adfwi_elastic_moduli_*, adfwi_elastic_pad_* through the C ABI) against
the fixtures of the unmodified reference (planes bit-identical, model-level gradients from the stored plane-level gradients), the
numpy oracle on a model-sized grid and the eager torch chain it replaces."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PLANES = ("C11", "C13", "C33", "C55", "bx", "bz")
PARAMS = ("vp", "vs", "rho", "eps", "delta")


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("name", ["elastic_pml_o4_fs", "elastic_gerjan_o6_nofs", "elastic_pml_o6_nofs"])
def test_fused_parameterisation_matches_reference(golden_dir, name):
    from adfwi_b200.model import thomsen_to_staggered_planes
    g = np.load(f"{golden_dir}/{name}.npz")
    leaves = {k: torch.tensor(g[k], device="cuda:0", requires_grad=True) for k in PARAMS}
    out = thomsen_to_staggered_planes(*[leaves[k] for k in PARAMS])
    for k, t in zip(PLANES, out):
        got = t.detach().cpu().numpy()
        if k == "C13":      # the fixture ran on the CPU, whose vectorised torch sqrt is 1 ulp off in places (a few ulp after the subtraction of C44); CUDA's sqrtf is correctly rounded
            assert np.abs(got.view(np.int32) - g["in_" + k].view(np.int32)).max() <= 4, k
        else:
            assert np.array_equal(got, g["in_" + k]), k
    for tag in ("stress", "vel"):
        for v in leaves.values():
            v.grad = None
        torch.autograd.backward(list(out), [torch.tensor(g[f"g_{k}_{tag}"], device="cuda:0") for k in PLANES], retain_graph=True)
        for k in PARAMS:
            assert rel_l2(leaves[k].grad.cpu().numpy(), g[f"g_{k}_{tag}"]) < 2e-5, (tag, k)


def test_fused_parameterisation_matches_oracle_and_eager_chain_on_a_model_sized_grid():
    from adfwi_b200 import synthetic as syn
    from adfwi_b200.model import FusedElasticGridModel
    from oracle import parameters_oracle as PO
    rng = np.random.default_rng(11)
    nz, nx = 350, 1700
    vp = syn.marmousi_like_vp(nz, nx); vs = (vp / 1.8).astype(np.float32); rho = syn.gardner_rho(vp)
    eps = (0.05 + 0.1 * rng.random((nz, nx))).astype(np.float32); delta = (-0.05 + 0.1 * rng.random((nz, nx))).astype(np.float32)
    ref = PO.planes(vp, vs, rho, eps, delta)
    models = [cls(vp, vs, rho, eps=eps, delta=delta, requires_grad=PARAMS, device="cuda:0") for cls in (FusedElasticGridModel, syn.ElasticGridModel)]
    W = {k: rng.standard_normal(ref[k].shape).astype(np.float32) for k in PLANES}
    grads = []
    for m in models:
        m.forward()
        got = dict(C11=m.CC[0], C13=m.CC[2], C33=m.CC[11], C55=m.CC[18], bx=m.bx, bz=m.bz)
        for k in PLANES:
            assert np.array_equal(got[k].detach().cpu().numpy(), ref[k]), (type(m).__name__, k)
        sum((got[k] * torch.tensor(W[k], device="cuda:0")).sum() for k in PLANES).backward()
        grads.append({k: getattr(m, k).grad.cpu().numpy() for k in PARAMS})
    want = PO.planes_T(vp, vs, rho, eps, delta, W)
    for k, w in zip(PARAMS, want):
        assert rel_l2(grads[0][k], w) < 2e-5, k
        assert rel_l2(grads[0][k], grads[1][k]) < 2e-5, k


@pytest.mark.parametrize("fs", [True, False])
@pytest.mark.parametrize("order", [4, 6])
def test_fused_padding_matches_the_eager_pads(fs, order):
    from adfwi_b200.propagator import elastic_kernels as ek
    torch.manual_seed(0)
    nz, nx, nabc = 37, 53, 9
    NN = order // 2
    nxp = nx + 2 * nabc
    nzp = nz + (nabc + NN if fs else 2 * nabc + NN)
    shapes = [(nz, nx), (nz, nx), (nz, nx), (nz - 2, nx - 2), (nz, nx - 1), (nz - 1, nx)]
    a = [torch.randn(s, device="cuda:0", requires_grad=True) for s in shapes]
    b = [t.detach().clone().requires_grad_(True) for t in a]
    fused = ek._PadPlanes.apply(*a, nz, nx, nzp, nxp, nabc, NN if fs else NN + nabc)
    eager = [ek.full_plane(t, nzp, nxp, nabc, NN, fs) for t in b]
    W = [torch.randn(nzp, nxp, device="cuda:0") for _ in range(6)]
    for f, e in zip(fused, eager):
        assert torch.equal(f, e)
    sum((f * w).sum() for f, w in zip(fused, W)).backward()
    sum((e * w).sum() for e, w in zip(eager, W)).backward()
    for x, y in zip(a, b):
        assert rel_l2(x.grad.cpu().numpy(), y.grad.cpu().numpy()) < 1e-6
