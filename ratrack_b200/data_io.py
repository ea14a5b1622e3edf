"""Input / output side of the hot path: View-of-Delft radar `.bin` records in, pinned
host batches for the engine, tracking result lines out.  Host code only.

* `read_radar_bin`      -- the reference's `FrameDataLoader.get_radar_scan` (src/vod/frame/data_loader.py:164-180):
                           float32 records of 7 columns [x, y, z, RCS, v_r, v_r_compensated, time].
* `frame_pair_inputs`   -- the slicing `main_utils.epoch` does before the model call (src/main_utils.py:76-79):
                           pc = columns 0:3 as (1,3,N), features = columns 3:5 as (1,2,N).
* `PaddedBatcher`       -- the batcher for real data: frame pairs of ANY point counts (real VoD frames carry 240-350
                           points each) go into one pinned batch, zero-padded to the largest count, together with the
                           per-cloud counts `npts1` / `npts2`.  Padding is not neutral for a point-cloud network -- a
                           sentinel point would change furthest-point sampling, ball-query slot filling and the kNN sets
                           -- so the convention is that padded points DO NOT EXIST: the engine's variable-size entry
                           (`rt_backbone_forward_varlen`, `Track4DBackbone.backbone(..., npts1=, npts2=)`) restricts every
                           search to a cloud's own points and samples every cloud with the tie-break of its own size, so a
                           pair's results are bit-identical to running it alone; output columns of padded points are zero.
* `PinnedBatcher`       -- the round-1 batcher: groups frame pairs BY POINT COUNT (one N per batch, no counts needed).
* `format_result_line`  -- one line of the reference's per-frame result file (src/main_utils.py:165-184).
"""
from collections import defaultdict

import numpy as np
import torch


def read_radar_bin(path):
    """-> (N,7) float32 array.  reference: np.fromfile(radar_file, dtype=np.float32).reshape(-1, 7)"""
    scan = np.fromfile(path, dtype=np.float32)
    if scan.size % 7:
        raise ValueError(f"{path}: {scan.size} floats is not a whole number of 7-column radar records")
    return scan.reshape(-1, 7)


def frame_pair_inputs(frame1, frame2):
    """two (N,7) records arrays of one point count -> pc1, pc2 (1,3,N), ft1, ft2 (1,2,N) float32 numpy, C-contiguous"""
    f1, f2 = np.asarray(frame1, np.float32), np.asarray(frame2, np.float32)
    if f1.shape[0] != f2.shape[0]:
        raise ValueError("the two frames of a pair must have the same number of points (the reference's loader resamples them)")

    def cols(f, a, b):
        return np.ascontiguousarray(f[:, a:b].T[None])
    return cols(f1, 0, 3), cols(f2, 0, 3), cols(f1, 3, 5), cols(f2, 3, 5)


class PinnedBatcher:
    """Collect frame pairs, hand out pinned batches of ONE point count.

        b = PinnedBatcher(batch=32)
        b.add(key, frame1, frame2)                 # (N,7) arrays; `key` comes back with the batch
        for keys, pc1, pc2, ft1, ft2 in b.ready(): # full batches (pinned torch tensors, reused buffers)
        for keys, ... in b.flush():                # whatever is left, as partial batches
    """

    def __init__(self, batch=32, pin=True):
        self.batch = int(batch)
        self.pin = pin and torch.cuda.is_available()
        self._pending = defaultdict(list)          # N -> [(key, pc1, pc2, ft1, ft2)]
        self._buffers = {}                         # (B, N) -> 4 pinned tensors

    def add(self, key, frame1, frame2):
        pc1, pc2, ft1, ft2 = frame_pair_inputs(frame1, frame2)
        self._pending[pc1.shape[2]].append((key, pc1, pc2, ft1, ft2))

    def _emit(self, n, items):
        b = len(items)
        if (b, n) not in self._buffers:
            mk = lambda c: (torch.empty(b, c, n).pin_memory() if self.pin else torch.empty(b, c, n))  # noqa: E731
            self._buffers[(b, n)] = (mk(3), mk(3), mk(2), mk(2))
        bufs = self._buffers[(b, n)]
        for i, it in enumerate(items):
            for buf, arr in zip(bufs, it[1:]):
                buf[i].copy_(torch.from_numpy(arr[0]))
        return ([it[0] for it in items],) + bufs

    def ready(self):
        for n in list(self._pending):
            q = self._pending[n]
            while len(q) >= self.batch:
                items, self._pending[n] = q[:self.batch], q[self.batch:]
                q = self._pending[n]
                yield self._emit(n, items)

    def flush(self):
        yield from self.ready()
        for n in list(self._pending):
            if self._pending[n]:
                items, self._pending[n] = self._pending[n], []
                yield self._emit(n, items)


class PaddedBatcher:
    """Collect frame pairs of any sizes, hand out pinned zero-padded batches with their point counts.

        b = PaddedBatcher(batch=32)
        b.add(key, frame1, frame2)                                # (N1,7), (N2,7) arrays: the two frames may differ in size
        for keys, pc1, pc2, ft1, ft2, npts1, npts2 in b.ready():  # (B,3,Nmax) / (B,2,Nmax) pinned tensors, (B,) int32 counts
        for ... in b.flush():                                     # whatever is left, as a partial batch
    """

    def __init__(self, batch=32, pin=True, multiple=4):
        self.batch, self.multiple = int(batch), int(multiple)
        self.pin = pin and torch.cuda.is_available()
        self._pending = []

    def add(self, key, frame1, frame2):
        f1, f2 = np.asarray(frame1, np.float32), np.asarray(frame2, np.float32)
        self._pending.append((key, f1, f2))

    def _emit(self, items):
        b = len(items)
        nmax = max(max(f1.shape[0], f2.shape[0]) for _, f1, f2 in items)
        nmax = (nmax + self.multiple - 1) // self.multiple * self.multiple      # 16-byte aligned clouds for the bulk-copy engine
        mk = lambda c: (torch.zeros(b, c, nmax).pin_memory() if self.pin else torch.zeros(b, c, nmax))  # noqa: E731
        pc1, pc2, ft1, ft2 = mk(3), mk(3), mk(2), mk(2)
        n1 = torch.tensor([f1.shape[0] for _, f1, _ in items], dtype=torch.int32)
        n2 = torch.tensor([f2.shape[0] for _, _, f2 in items], dtype=torch.int32)
        for i, (_, f1, f2) in enumerate(items):
            pc1[i, :, :f1.shape[0]] = torch.from_numpy(f1[:, 0:3].T.copy())
            ft1[i, :, :f1.shape[0]] = torch.from_numpy(f1[:, 3:5].T.copy())
            pc2[i, :, :f2.shape[0]] = torch.from_numpy(f2[:, 0:3].T.copy())
            ft2[i, :, :f2.shape[0]] = torch.from_numpy(f2[:, 3:5].T.copy())
        return [it[0] for it in items], pc1, pc2, ft1, ft2, n1, n2

    def ready(self):
        while len(self._pending) >= self.batch:
            items, self._pending = self._pending[:self.batch], self._pending[self.batch:]
            yield self._emit(items)

    def flush(self):
        yield from self.ready()
        if self._pending:
            items, self._pending = self._pending, []
            yield self._emit(items)


def format_result_line(obj_id, conf, obj):
    """obj (1,C>=6,k) tensor of one tracked object -> the reference's result line
    'NA 1 -1 -1 <conf> <id> x y z x y z ...' with the points' pc1 coordinates (feature rows 3:6).  main_utils.py:165-184"""
    parts = ["NA", "1", "-1", "-1", str(float(conf)), str(obj_id)]
    xyz = obj[0, 3:6, :].detach().cpu().numpy()
    for i in range(xyz.shape[1]):
        parts += [str(float(xyz[0, i])), str(float(xyz[1, i])), str(float(xyz[2, i]))]
    return " ".join(parts) + "\n"
