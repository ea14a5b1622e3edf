"""Input / output side: radar record reader, model-input slicing, batching by point count,
result lines.  The three real VoD frames the reference ships are in tests/golden/pointnet2_vod_frames.npz (xyz only)."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from ratrack_b200 import data_io


def _frames():
    g = np.load(os.path.join(GOLDEN, "pointnet2_vod_frames.npz"))
    out = []
    for tag in sorted(k[4:] for k in g.files if k.startswith("xyz_")):
        xyz = g[f"xyz_{tag}"]
        rec = np.zeros((xyz.shape[0], 7), np.float32)
        rec[:, 0:3] = xyz
        rec[:, 3] = -13.0
        rec[:, 4] = np.linspace(-3, 1, xyz.shape[0], dtype=np.float32)
        out.append(rec)
    return out


def test_read_radar_bin_round_trip(tmp_path):
    rec = _frames()[0]
    p = tmp_path / "00549.bin"
    rec.tofile(p)                                                       # the reference reads exactly this layout
    got = data_io.read_radar_bin(str(p))
    assert got.dtype == np.float32 and got.shape == rec.shape and np.array_equal(got, rec)
    (tmp_path / "bad.bin").write_bytes(b"\0" * 40)
    try:
        data_io.read_radar_bin(str(tmp_path / "bad.bin"))
        assert False
    except ValueError:
        pass


def test_frame_pair_inputs_layout():
    a = _frames()[0]
    pc1, pc2, ft1, ft2 = data_io.frame_pair_inputs(a, a[::-1])
    n = a.shape[0]
    assert pc1.shape == (1, 3, n) and ft1.shape == (1, 2, n) and pc1.flags["C_CONTIGUOUS"]
    assert np.array_equal(pc1[0].T, a[:, 0:3]) and np.array_equal(ft2[0].T, a[::-1][:, 3:5])    # main_utils.py:76-79


def test_batcher_groups_by_point_count_without_padding():
    fr = _frames()                                                      # N = 322, 352, 242
    b = data_io.PinnedBatcher(batch=2, pin=False)
    order = [0, 1, 0, 2, 1, 0]
    for k, i in enumerate(order):
        b.add(k, fr[i], fr[i])
    full = list(b.ready())
    assert sorted(len(x[0]) for x in full) == [2, 2]
    for keys, pc1, pc2, ft1, ft2 in full:
        n = pc1.shape[2]
        assert pc1.shape == (2, 3, n) and ft1.shape == (2, 2, n)
        for row, key in enumerate(keys):
            src = fr[order[key]]
            assert src.shape[0] == n and torch.equal(pc1[row], torch.from_numpy(src[:, 0:3].T.copy()))
    rest = list(b.flush())
    assert sorted(len(x[0]) for x in rest) == [1, 1] and sorted(k for x in rest for k in x[0]) == [3, 5]
    assert list(b.flush()) == []


def test_result_line_format():
    obj = torch.zeros(1, 139, 2)
    obj[0, 3:6, 0] = torch.tensor([1.0, 2.0, 3.0])
    obj[0, 3:6, 1] = torch.tensor([4.0, 5.0, 6.5])
    assert data_io.format_result_line(7, 0.25, obj) == "NA 1 -1 -1 0.25 7 1.0 2.0 3.0 4.0 5.0 6.5\n"


def test_padded_batcher_keeps_every_frame_at_its_own_size():
    """Frames of different point counts share one zero-padded batch; the counts travel with it."""
    f = _frames()                                                        # 322, 352, 242 points
    b = data_io.PaddedBatcher(batch=3, pin=False)
    b.add("a", f[0], f[1])
    b.add("b", f[1], f[2])
    assert list(b.ready()) == []
    b.add("c", f[2], f[0])
    (keys, pc1, pc2, ft1, ft2, n1, n2), = list(b.ready())
    assert keys == ["a", "b", "c"] and n1.tolist() == [322, 352, 242] and n2.tolist() == [352, 242, 322]
    assert tuple(pc1.shape) == (3, 3, 352) and tuple(ft2.shape) == (3, 2, 352) and n1.dtype == torch.int32
    assert np.array_equal(pc1[2, :, :242].numpy(), f[2][:, 0:3].T) and float(pc1[2, :, 242:].abs().max()) == 0.0
    assert np.array_equal(ft2[1, :, :242].numpy(), f[2][:, 3:5].T) and float(ft2[1, :, 242:].abs().max()) == 0.0
    b.add("d", f[0], f[0])
    (keys, pc1, *_rest), = list(b.flush())
    assert keys == ["d"] and tuple(pc1.shape) == (1, 3, 324)             # padded to a multiple of 4 points
