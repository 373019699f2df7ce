"""Shot-sharded observed-record store (adfwi_b200/fwi/records.py): conversion from the reference's pickled-dict .npz
(ADFWI/survey/data.py:63-99), shard views, batch access."""
import numpy as np
import pytest

from adfwi_b200 import distributed as D
from adfwi_b200.fwi import records as R


def _reference_style_npz(path, data, nt, dt):
    ns, _, nr = next(iter(data.values())).shape
    np.savez(path, data=data, src_loc=np.zeros((ns, 2)), rcv_loc=np.zeros((nr, 2)), src_num=ns, rcv_num=nr, rcv_type=np.array(["pr"] * nr),
             src_type=np.array(["mt"] * ns), t=np.arange(nt) * dt, nt=nt, dt=dt)


def test_convert_and_shard(tmp_path):
    rng = np.random.default_rng(0)
    ns, nt, nr = 7, 50, 11
    data = {"p": rng.standard_normal((ns, nt, nr)).astype(np.float32), "u": rng.standard_normal((ns, nt, nr)).astype(np.float32)}
    src = str(tmp_path / "obs_data.npz")
    _reference_style_npz(src, data, nt, 1e-3)
    out = R.convert(src, str(tmp_path / "store"))
    parts = []
    for rank in range(3):
        lo, hi = D.shard_shots(ns, rank, 3)
        st = R.ShotRecordStore(out, shard=(lo, hi))
        assert len(st) == hi - lo and st.components == ["p", "u"]
        got = st.batch("p", np.arange(len(st)))
        assert np.array_equal(got.numpy(), data["p"][lo:hi])
        parts.append(st.host("u", np.arange(len(st))))
        if len(st) > 1:
            assert np.array_equal(st.host("u", [len(st) - 1, 0]), data["u"][[hi - 1, lo]])
        with pytest.raises(IndexError):
            st.host("p", [len(st)])
    assert np.array_equal(np.concatenate(parts), data["u"])
    n = R.ShotRecordStore(out, normalize=True).batch("p", [0, 1]).numpy()
    assert np.allclose(np.abs(n).max(axis=1), 1.0)


def test_reference_seismicdata_file_is_readable(tmp_path):
    """A file written by the unmodified reference's SeismicData.save converts as is."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    ref_loader.load()
    from ADFWI.survey import Source, Receiver, Survey, SeismicData
    nt, dt = 20, 1e-3
    s = Source(nt=nt, dt=dt, f0=10.0); s.add_sources(src_x=np.array([1, 3]), src_z=np.array([1, 1]), src_wavelet=np.zeros(nt), src_type="mt", src_mt=np.eye(3))
    r = Receiver(nt=nt, dt=dt); r.add_receivers(rcv_x=np.arange(5), rcv_z=np.ones(5, dtype=int), rcv_type="pr")
    sd = SeismicData(Survey(source=s, receiver=r))
    rec = np.random.default_rng(1).standard_normal((2, nt, 5)).astype(np.float32)
    import torch
    sd.record_data({"p": torch.tensor(rec)})
    path = str(tmp_path / "obs.npz")
    sd.save(path)
    st = R.ShotRecordStore(R.convert(path, str(tmp_path / "s")), shard=(1, 2))
    assert np.array_equal(st.host("p", [0]), rec[1:2])
