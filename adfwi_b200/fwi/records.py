"""Observed-record store that streams per shot shard (second half).

The reference keeps observed data as ONE ``.npz`` holding a pickled dict of ``(n_shots, nt, n_receivers)`` arrays
(``SeismicData.save`` / ``.load``, ADFWI/survey/data.py:63-99): every process that wants any shot unpickles all of them.  With shots
sharded over ranks (``adfwi_b200.distributed.shard_shots``) each rank only needs its own block, so:

  * :func:`convert` rewrites the reference file once into one plain ``.npy`` per record component plus a small JSON header -- a layout
    that can be memory-mapped;
  * :class:`ShotRecordStore` maps the files, exposes the rank's contiguous shard and hands shot batches to the device through a
    double-buffered PINNED staging area and a non-blocking copy, so the pages of other ranks' shots are never read and the host-to-device
    copy of batch k+1 overlaps the propagation of batch k."""
import json
import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

HEADER = "records.json"


def convert(src, out_dir: str, components: Optional[Sequence[str]] = None) -> str:
    """``src``: path of a reference ``obs_data.npz`` (pickled dict under ``data``) or a dict {component: (ns, nt, nr) array}."""
    if isinstance(src, str):
        z = np.load(src, allow_pickle=True)
        data = z["data"].item()
        meta = {k: (z[k].tolist() if k in z.files else None) for k in ("src_num", "rcv_num", "nt", "dt")}
    else:
        data, meta = dict(src), {}
    os.makedirs(out_dir, exist_ok=True)
    comps = list(components or data.keys())
    shape = None
    for c in comps:
        a = np.ascontiguousarray(np.asarray(data[c], dtype=np.float32))
        if a.ndim != 3:
            raise ValueError(f"record component {c!r} must be (n_shots, nt, n_receivers)")
        shape = shape or a.shape
        if a.shape != shape:
            raise ValueError("all record components must share one shape")
        np.save(os.path.join(out_dir, f"{c}.npy"), a)
    with open(os.path.join(out_dir, HEADER), "w") as f:
        json.dump({"components": comps, "shape": list(shape), "dtype": "float32", **{k: v for k, v in meta.items() if v is not None}}, f)
    return out_dir


class ShotRecordStore:
    def __init__(self, path: str, shard: Optional[Tuple[int, int]] = None, device=None, normalize: bool = False):
        with open(os.path.join(path, HEADER)) as f:
            self.header = json.load(f)
        self.components = list(self.header["components"])
        self.ns, self.nt, self.nr = self.header["shape"]
        self.lo, self.hi = shard if shard is not None else (0, self.ns)
        if not (0 <= self.lo <= self.hi <= self.ns):
            raise IndexError("shot shard out of range")
        self._maps: Dict[str, np.ndarray] = {c: np.load(os.path.join(path, f"{c}.npy"), mmap_mode="r") for c in self.components}
        self.device = torch.device(device) if device is not None else None
        self.normalize = normalize          # per-trace max-abs normalisation, as the reference drivers apply to obs once (acoustic_fwi.py:68-70)
        self._stage = {}                    # (component, slot) -> pinned staging tensor
        self._slot = 0

    def __len__(self):
        return self.hi - self.lo

    def host(self, component: str, positions) -> np.ndarray:
        """Records of the shard's shots ``positions`` (indices relative to the shard) as a numpy array (copied out of the map)."""
        pos = np.asarray(positions)
        if pos.size and (pos.min() < 0 or pos.max() >= len(self)):
            raise IndexError("shot position outside this rank's shard")
        m = self._maps[component]
        if pos.size and np.all(np.diff(pos) == 1):
            return np.array(m[self.lo + pos[0]: self.lo + pos[-1] + 1])
        return np.array(m[self.lo + pos])

    def batch(self, component: str, positions) -> torch.Tensor:
        """The same on ``device``: staged through pinned memory, copied without blocking the host (two slots alternate, so the copy of
        the next batch may be issued while the previous one is still in flight)."""
        a = self.host(component, positions)
        if self.device is None or self.device.type != "cuda":
            t = torch.from_numpy(a)
        else:
            key = (component, self._slot, a.shape)
            st = self._stage.get(key)
            if st is None:
                st = self._stage[key] = torch.empty(a.shape, dtype=torch.float32).pin_memory()
            st.copy_(torch.from_numpy(a))
            t = st.to(self.device, non_blocking=True)
            self._slot ^= 1
        if self.normalize:
            t = t / torch.max(torch.abs(t), dim=1, keepdim=True).values
        return t

    def loader(self, components: Sequence[str]):
        """``obs_loader`` for :func:`adfwi_b200.fwi.acoustic_gradient` / ``elastic_gradient``."""
        if len(components) == 1:
            return lambda pos: self.batch(components[0], pos)
        return lambda pos: {c: self.batch(c, pos) for c in components}
