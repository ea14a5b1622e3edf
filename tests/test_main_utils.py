"""CPU: the caller of the hot path (ratrack_b200/main_utils.py) -- frame order of the reference's dataset class
(src/dataset_classes/track_vod_3d.py:49-122) and the evaluation branch of its epoch loop (src/main_utils.py:44-185) --
driven with a recording stand-in for the network."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from ratrack_b200 import main_utils, metrics


def _write_frames(d, numbers, rng, sizes=None):
    os.makedirs(d, exist_ok=True)
    out = {}
    for i, n in enumerate(numbers):
        rec = rng.normal(size=(sizes[i] if sizes else 20 + n % 5, 7)).astype(np.float32)
        rec.tofile(os.path.join(d, str(n).zfill(5) + ".bin"))
        out[n] = rec
    return out


def _write_clip(d, name, numbers):
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, name + ".txt"), "w") as f:
        f.write("\n".join(str(n).zfill(5) for n in numbers) + "\n")


def test_clip_frame_pairs_follow_the_reference_order(tmp_path):
    rng = np.random.default_rng(0)
    radar, clips = str(tmp_path / "radar"), str(tmp_path / "clips")
    frames = _write_frames(radar, [9, 10, 11, 12, 14, 15, 29, 30, 31, 32], rng)    # 13 is missing
    with open(os.path.join(radar, "00033.bin"), "wb") as f:
        f.write(b"\0" * 40)                                                        # not a whole number of records
    _write_clip(clips, "delft_1", range(10, 16))
    _write_clip(clips, "delft_10", range(30, 35))                                  # 34 is missing too
    got = list(main_utils.clip_frame_pairs(radar, clips, ["delft_1", "delft_10"]))
    # (first cloud = frame current + 1, second = frame current, index = current + 1); pairs touching 13 / 33 / 34 are skipped --
    # as first, second or PREVIOUS frame (the reference loads frame current - 1 too): (15, 14) goes with them
    assert [(g[2], g[3], g[4]) for g in got] == [(11, "delft_1", True), (12, "delft_1", False),
                                                 (31, "delft_10", True), (32, "delft_10", False)]
    loose = list(main_utils.clip_frame_pairs(radar, clips, ["delft_1"], need_previous=False))
    assert [g[2] for g in loose] == [11, 12, 15]
    for a, b, index, _, _ in got:
        assert np.array_equal(a, frames[index]) and np.array_equal(b, frames[index - 1])


def test_new_sequence_flag_survives_skipped_leading_frames(tmp_path):
    rng = np.random.default_rng(1)
    radar, clips = str(tmp_path / "radar"), str(tmp_path / "clips")
    _write_frames(radar, [5, 6, 7], rng)                                            # the clip starts at 3: 3, 4 do not exist
    _write_clip(clips, "c", range(3, 8))
    got = [(g[2], g[4]) for g in main_utils.clip_frame_pairs(radar, clips, ["c"], need_previous=False)]
    assert got == [(6, True), (7, False)]
    # the reference also needs frame current - 1 (its lidar sweep): the pair (6, 5) goes too, 4 does not exist
    assert [(g[2], g[4]) for g in main_utils.clip_frame_pairs(radar, clips, ["c"])] == [(7, True)]


def test_pair_tensors_equal_and_unequal_sizes():
    rng = np.random.default_rng(2)
    a, b = rng.normal(size=(22, 7)).astype(np.float32), rng.normal(size=(22, 7)).astype(np.float32)
    pc1, pc2, ft1, ft2, n1, n2 = main_utils.pair_tensors(a, b)
    assert n1 is None and n2 is None and pc1.shape == (1, 3, 22) and ft2.shape == (1, 2, 22)
    assert np.array_equal(pc1[0].numpy(), a[:, :3].T) and np.array_equal(ft2[0].numpy(), b[:, 3:5].T)   # main_utils.py:76-79
    c = rng.normal(size=(17, 7)).astype(np.float32)
    pc1, pc2, ft1, ft2, n1, n2 = main_utils.pair_tensors(a, c)
    assert n1.tolist() == [22] and n2.tolist() == [17] and pc1.shape == (1, 3, 24) and pc2.shape == (1, 3, 24)
    assert np.array_equal(pc2[0, :, :17].numpy(), c[:, :3].T) and float(pc2[0, :, 17:].abs().max()) == 0.0
    assert np.array_equal(ft1[0, :, :22].numpy(), a[:, 3:5].T) and float(ft1[0, :, 22:].abs().max()) == 0.0


class _RecordingNet:
    """Stands in for Track4D: records what the loop passes and returns the reference's 10-tuple (track4d.py:65)."""

    def __init__(self):
        self.calls = []
        self.next_id = 0

    def __call__(self, pc1, pc2, ft1, ft2, h, objects_prev, npts1=None, npts2=None):
        n = pc1.shape[-1] if npts1 is None else int(npts1[0])
        self.calls.append({"h": h, "prev": dict(objects_prev), "n": n, "padded": pc1.shape[-1], "npts": (npts1, npts2),
                           "shapes": (tuple(pc1.shape), tuple(pc2.shape), tuple(ft1.shape), tuple(ft2.shape))})
        h_new = torch.full((5, 1, 128), float(len(self.calls)), requires_grad=True)
        pc1 = pc1[..., :n]
        flow = torch.full_like(pc1, 0.25)
        cls = (torch.arange(n) % 2).float().reshape(1, n)                          # every second point moves
        objects = {}
        for _ in range(2):                                                         # two objects per frame, fresh ids
            objects[self.next_id] = torch.cat([pc1 + flow, pc1, flow], 1)[:, :, :3].clone().requires_grad_(True)
            self.next_id += 1
        confs = torch.tensor([0.75, 0.5])
        return h_new, pc1 + flow, cls, [], None, None, confs, objects, dict(), []


def test_eval_epoch_carries_state_writes_results_and_accumulates_metrics(tmp_path):
    rng = np.random.default_rng(3)
    radar, clips, results = str(tmp_path / "radar"), str(tmp_path / "clips"), str(tmp_path / "results")
    frames = _write_frames(radar, [1, 2, 3, 4, 8, 9], rng, sizes=[24, 24, 19, 24, 16, 16])
    _write_clip(clips, "a", range(1, 5))
    _write_clip(clips, "b", range(8, 10))
    net = _RecordingNet()

    def gt_fn(clip, index, pc1):
        if index == 3:                                                             # a frame without labels (the reference's `continue`)
            return None
        return torch.full_like(pc1, 0.5), (torch.arange(pc1.shape[-1]) % 2).float().reshape(1, -1)

    out = main_utils.eval_epoch(net, main_utils.clip_frame_pairs(radar, clips, ["a", "b"], need_previous=False), results_dir=results, gt_fn=gt_fn)
    assert out["frames"] == 4 and out["objects"] == 8 and out["examples"] == 3
    c = net.calls
    # state: reset at each new sequence, carried (detached) inside one
    assert c[0]["h"] is None and c[0]["prev"] == {}
    assert float(c[1]["h"].flatten()[0]) == 1.0 and not c[1]["h"].requires_grad and sorted(c[1]["prev"]) == [0, 1]
    assert not any(v.requires_grad for v in c[1]["prev"].values())
    assert sorted(c[2]["prev"]) == [2, 3]
    assert c[3]["h"] is None and c[3]["prev"] == {}                                # clip "b"
    # sizes: (2,1) equal; (3,2) and (4,3) unequal -> padded to a multiple of 4 with counts
    assert c[0]["npts"] == (None, None) and c[0]["shapes"][0] == (1, 3, 24)
    assert c[1]["npts"][0].tolist() == [19] and c[1]["npts"][1].tolist() == [24] and c[1]["padded"] == 24
    assert c[2]["npts"][0].tolist() == [24] and c[2]["npts"][1].tolist() == [19]
    # result files: <results>/<clip>/<index:05d>.txt, 'NA 1 -1 -1 conf id x y z ...' per object (main_utils.py:167-184)
    assert sorted(os.listdir(os.path.join(results, "a"))) == ["00002.txt", "00003.txt", "00004.txt"]
    assert os.listdir(os.path.join(results, "b")) == ["00009.txt"]
    lines = open(os.path.join(results, "a", "00003.txt")).read().splitlines()
    assert len(lines) == 2
    f = lines[1].split(" ")
    assert f[:4] == ["NA", "1", "-1", "-1"] and float(f[4]) == 0.5 and f[5] == "3" and len(f) == 6 + 3 * 3
    assert [float(x) for x in f[6:9]] == [float(v) for v in frames[3][0, :3]]      # rows 3:6 of an object = pc1 coordinates
    # metrics: sums over the frames with ground truth, equal to calling the metric functions frame by frame
    want_flow, want_seg = {}, {}
    for index in (2, 4, 9):
        pc1 = torch.from_numpy(np.ascontiguousarray(frames[index][:, :3].T[None]))
        n = pc1.shape[-1]
        cls = (torch.arange(n) % 2).float().reshape(1, n)
        metrics.accumulate(want_flow, metrics.eval_scene_flow(pc1, pc1 + 0.25, torch.full_like(pc1, 0.5), cls))
        metrics.accumulate(want_seg, metrics.eval_motion_seg((cls > 0.5).float(), cls))
    assert out["flow"] == pytest.approx(metrics.as_floats(want_flow), nan_ok=True)
    assert out["seg"] == pytest.approx(metrics.as_floats(want_seg))
    assert out["seg"]["acc"] == pytest.approx(3.0)


def test_track4d_forward_drops_the_padded_columns_before_tracking():
    """Track4D.forward(npts1=, npts2=): the variable-size backbone entry, then the reference's tail on the unpadded pair."""
    from ratrack_b200.track4d import Track4D

    class A:
        npoints = 512
        min_obj_points = 2

    seen = {}

    class T(Track4D):
        def backbone(self, pc1, pc2, feature1, feature2, h, npts1=None, npts2=None):
            seen["npts"] = (npts1, npts2)
            B, _, N = pc1.shape
            z = lambda c: torch.zeros(B, c, N)                                      # noqa: E731
            return z(3), torch.zeros(5, B, 128), torch.zeros(B, N), z(256), z(256), z(256), z(128)

        def track(self, pc1, feature1, backbone_out, objects_prev):
            seen["track"] = (tuple(pc1.shape), tuple(feature1.shape), [None if o is None else tuple(o.shape) for o in backbone_out])
            return "tail"

    net = T(A())
    n1, n2 = torch.tensor([19], dtype=torch.int32), torch.tensor([22], dtype=torch.int32)
    r = net(torch.zeros(1, 3, 24), torch.zeros(1, 3, 24), torch.zeros(1, 2, 24), torch.zeros(1, 2, 24), None, {}, npts1=n1, npts2=n2)
    assert r == "tail" and seen["npts"] == (n1, n2)
    pc1_s, ft1_s, outs = seen["track"]
    assert pc1_s == (1, 3, 19) and ft1_s == (1, 2, 19)
    assert outs == [(1, 3, 19), (5, 1, 128), (1, 19), (1, 256, 19), (1, 256, 19), (1, 256, 22), (1, 128, 19)]
    net(torch.zeros(1, 3, 24), torch.zeros(1, 3, 24), torch.zeros(1, 2, 24), torch.zeros(1, 2, 24), None, {})
    assert seen["npts"] == (None, None) and seen["track"][0] == (1, 3, 24)
    with pytest.raises(ValueError):
        net(torch.zeros(2, 3, 24), torch.zeros(2, 3, 24), torch.zeros(2, 2, 24), torch.zeros(2, 2, 24), None, {}, npts1=n1, npts2=n2)


def test_clips_are_dealt_whole_and_summaries_add_up_without_a_process_group():
    clips = ["delft_1", "delft_10", "delft_14", "delft_22", "delft_7"]
    parts = [main_utils.clips_of_rank(clips, r, 3) for r in range(3)]
    assert parts == [["delft_1", "delft_22"], ["delft_10", "delft_7"], ["delft_14"]]
    assert sorted(sum(parts, [])) == sorted(clips)
    with pytest.raises(ValueError):
        main_utils.clips_of_rank(clips, 3, 3)
    s = {"frames": 4, "objects": 8, "examples": 3, "flow": {k: float(i) for i, k in enumerate(main_utils.FLOW_KEYS)},
         "seg": {"acc": 3.0, "miou": 1.5, "sen": 2.0}}
    assert main_utils.reduce_summary(s) == s                                        # no process group: a copy
    empty = {"frames": 2, "objects": 0, "examples": 0, "flow": {}, "seg": {}}
    assert main_utils.reduce_summary(empty) == empty


def _gt(clip, index, pc1):
    return torch.full_like(pc1, 0.5), (torch.arange(pc1.shape[-1]) % 2).float().reshape(1, -1)


def _eval_worker(rank, world, port, root, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    clips = main_utils.clips_of_rank(["a", "b", "c"], rank, world)
    pairs = main_utils.clip_frame_pairs(os.path.join(root, "radar"), os.path.join(root, "clips"), clips, need_previous=False)
    local = main_utils.eval_epoch(_RecordingNet(), pairs, results_dir=os.path.join(root, "results"), gt_fn=_gt)
    total = main_utils.reduce_summary(local)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, local["frames"], total))


def test_two_rank_gloo_evaluation_over_whole_clips(tmp_path):
    """a sequence stays on one rank; the summed summary equals the single-process loop over all clips."""
    rng = np.random.default_rng(4)
    root = str(tmp_path)
    _write_frames(os.path.join(root, "radar"), list(range(1, 5)) + list(range(10, 13)) + list(range(20, 26)), rng)
    for name, r in (("a", range(1, 5)), ("b", range(10, 13)), ("c", range(20, 26))):
        _write_clip(os.path.join(root, "clips"), name, r)
    want = main_utils.eval_epoch(_RecordingNet(), main_utils.clip_frame_pairs(os.path.join(root, "radar"), os.path.join(root, "clips"),
                                                                             ["a", "b", "c"], need_previous=False), gt_fn=_gt)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, root, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [3 + 5, 2]                                        # rank 0: clips a, c; rank 1: clip b
    for _, _, total in res:
        assert total["frames"] == want["frames"] == 10 and total["objects"] == want["objects"] and total["examples"] == 10
        assert total["seg"] == pytest.approx(want["seg"])
        assert total["flow"] == pytest.approx(want["flow"], nan_ok=True)
    written = sorted(os.path.join(c, f) for c in os.listdir(os.path.join(root, "results")) for f in os.listdir(os.path.join(root, "results", c)))
    assert len(written) == 10 and written[0] == os.path.join("a", "00002.txt")


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree only exists in the dev container")
def test_frame_order_equals_the_unmodified_reference_dataset_class(tmp_path):
    """oracle/ref_frame_order.py drives the reference's own TrackingDataVOD.__getitem__ (file layer substituted, control flow
    untouched) over the same directory: same pairs, indices, clips and new-sequence flags -- including the pair it drops at
    the first frame of the data set (no frame -1) and around a missing frame."""
    import json
    import subprocess
    import sys

    from conftest import ROOT

    rng = np.random.default_rng(6)
    radar, clips = str(tmp_path / "radar"), str(tmp_path / "clips")
    numbers = [n for n in range(0, 8) if n != 4] + list(range(20, 24)) + list(range(40, 43)) + list(range(60, 64))
    frames = _write_frames(radar, numbers, rng)
    val = ["delft_1", "delft_10", "delft_14", "delft_22"]                           # track_vod_3d.py:34, args.eval = True
    for name, r in zip(val, (range(0, 8), range(20, 24), range(40, 43), range(60, 64))):
        _write_clip(clips, name, r)
    mine = list(main_utils.clip_frame_pairs(radar, clips, val))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_frame_order.py"), radar, str(len(mine))],
                       capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    ref = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(ref) == len(mine) and len(mine) >= 8
    for (a, b, index, clip, new_seq), want in zip(mine, ref):
        assert (index, clip, new_seq, a.shape[0], b.shape[0]) == (want["index"], want["clip"], want["new_seq"], want["n0"], want["n1"])
        assert float(np.float64(a[:, :5]).sum()) == pytest.approx(want["sum0"], rel=1e-12)
        assert float(np.float64(b[:, :5]).sum()) == pytest.approx(want["sum1"], rel=1e-12)
        assert np.array_equal(a, frames[index]) and np.array_equal(b, frames[index - 1])
    assert [m[2] for m in mine][:4] == [2, 3, 7, 22] and mine[0][4] and mine[3][4]   # (1,0), (4,3), (5,4), (6,5) dropped
