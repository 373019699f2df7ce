"""Host-side logic: shot sharding / batching, memory-scheme selection, padding helpers, synthetic
parameterisation (CPU, no compute through the CUDA library)."""
import ctypes as C

import numpy as np
import torch

from adfwi_b200 import distributed as D, fwi


def test_shard_shots_partitions_exactly():
    for n in (1, 7, 30, 240, 241):
        for w in (1, 2, 3, 8):
            parts = [D.shard_shots(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_shot_batches_match_reference_rule():
    # acoustic_fwi.py:136-138: ceil(n/b) batches, the last one takes the remainder
    assert [list(b) for b in fwi.shot_batches(5, 2)] == [[0, 1], [2, 3], [4]]
    assert [list(b) for b in fwi.shot_batches(4, None)] == [[0, 1, 2, 3]]
    assert [list(b) for b in fwi.shot_batches(3, 10)] == [[0, 1, 2]]


def test_ckpt_interval_choice_fits_budget():
    import emul_driver as E
    from adfwi_b200.propagator import acoustic_kernels as ak
    lib = E.emul_lib()
    d = ak.make_desc(450, 1800, 10, 4000, 1700, 50, True, 1e-3, 1, True, 0, False, 0)
    full = lib.adfwi_acoustic_workspace_bytes(C.byref(d))
    assert full > 10 * 4000 * 450 * 1800 * 4              # store-all keeps one plane per step per shot
    K, need = ak.choose_ckpt_interval(lib, d, full)
    assert K == 0 and need == full
    K, need = ak.choose_ckpt_interval(lib, d, full // 3)
    assert 0 < K < 4000 and need <= full // 3
    d.ckpt_interval = K
    assert lib.adfwi_acoustic_workspace_bytes(C.byref(d)) == need


def test_padding_helpers_match_oracle():
    from adfwi_b200.propagator import acoustic_kernels as ak, elastic_kernels as ek
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    a = rng.standard_normal((7, 9)).astype(np.float32)
    assert np.array_equal(ak.pad_replicate(torch.tensor(a), 3).numpy(), O.acoustic_pad(a, 3))
    for fs in (True, False):
        for shape in ((7, 9), (7, 8), (6, 9), (5, 7)):          # full, bx, bz, C55 ragged shapes of a 7x9 model
            b = rng.standard_normal(shape).astype(np.float32)
            nzp = 7 + (3 + 2 if fs else 6 + 2); nxp = 9 + 6
            got = ek.full_plane(torch.tensor(b), nzp, nxp, 3, 2, fs).numpy()
            assert np.array_equal(got, O.elastic_full_plane(b, nzp, nxp, 3, 2, fs))


def test_acoustic_coefficients_match_oracle_bits():
    from adfwi_b200.propagator import acoustic_kernels as ak
    from oracle import oracle as O
    rng = np.random.default_rng(1)
    vp = (2000 + 500 * rng.random((12, 14))).astype(np.float32)
    rho = (2000 + 100 * rng.random((12, 14))).astype(np.float32)
    damp = (50 * rng.random((20, 22))).astype(np.float32)
    for fs in (True, False):
        ref = O.acoustic_coefficients(vp, rho, damp, 1e-3, 10.0, 4, fs)
        got = ak.coefficient_planes(torch.tensor(vp), torch.tensor(rho), torch.tensor(damp), 1e-3, 10.0, 4, fs)
        for t, k in zip(got, ("alpha1", "alpha2", "kappa1", "kappa2", "kappa3")):
            assert np.array_equal(t.numpy(), ref[k]), k


def test_diff_coef_bits():
    from adfwi_b200.propagator import elastic_kernels as ek
    from oracle import oracle as O
    for NN in (2, 3):
        assert np.array_equal(np.array(ek.diff_coef(NN), dtype=np.float32), O.diff_coef(NN))
    assert abs(ek.diff_coef(3)[1] - (-25.0 / 384.0)) < 1e-8


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the driver's reference arm) runs on the host only: one JSON line with the keys the
    contract names, timed on the unmodified reference (when present) with the oracle port beside it."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "C1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=root, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gcell-updates/s" and d["value"] > 0
    from oracle import ref_loader
    # the unmodified reference when its package is present (this container / staged under baseline/_ref), else the C port
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["port"]["kind"] == "port" and d["port"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_host_cores_is_bounded_by_what_the_process_may_use():
    """bench.host_cores(): never more than the visible CPUs, the affinity mask or the cgroup quota, never less than one."""
    import importlib.util
    import os
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    n = bench.host_cores()
    assert 1 <= n <= (os.cpu_count() or 1)
    assert n <= len(os.sched_getaffinity(0))
