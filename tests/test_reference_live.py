"""Optional: checks against the LIVE reference when its tree is present (this container); skipped
on the GPU box, where only the committed golden fixtures travel."""
import numpy as np
import pytest

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


def test_diffcoef_bits_from_reference():
    ref_loader.load()
    from ADFWI.propagator import elastic_kernels as ek
    from oracle import oracle as O
    for NN in (2, 3):
        assert np.array_equal(ek.DiffCoef(NN, "s").numpy(), O.diff_coef(NN))


def test_fresh_acoustic_case_bit_identical():
    import torch
    ref_loader.load()
    from ADFWI.propagator import acoustic_kernels as ak
    from oracle import oracle as O
    rng = np.random.default_rng(123)
    nz, nx, nabc, nt = 21, 26, 6, 60
    vp = (1800 + 900 * rng.random((nz, nx))).astype(np.float32)
    rho = (1900 + 300 * rng.random((nz, nx))).astype(np.float32)
    damp = (40 * rng.random((nz + 2 * nabc, nx + 2 * nabc))).astype(np.float32)
    sx, sz = np.array([4, 20]), np.array([0, 9]); rx, rz = np.array([0, 13, 25]), np.array([0, 3, 20])
    sv = rng.standard_normal((2, nt)).astype(np.float32)
    for fs in (True, False):
        rec = ak.forward_kernel(nx, nz, 10.0, 10.0, nt, 1e-3, nabc, fs, torch.tensor(sx), torch.tensor(sz), 2, torch.tensor(sv),
                                torch.tensor(rx), torch.tensor(rz), 3, torch.tensor(damp), torch.tensor(vp), torch.tensor(rho),
                                checkpoint_segments=2)
        coef = O.acoustic_coefficients(vp, rho, damp, 1e-3, 10.0, nabc, fs)
        out = O.acoustic_run(coef, nabc, fs, 1e-3, sx, sz, sv, rx, rz)
        for k in "puw":
            assert np.array_equal(out[k], rec[k].numpy())
