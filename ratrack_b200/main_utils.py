"""The caller of the hot path: the evaluation branch of the reference's `epoch` loop (src/main_utils.py:44-185) and the
frame order of its dataset class (src/dataset_classes/track_vod_3d.py:49-122), restricted to what drives `Track4D.forward`.
Host code only.

* `clip_frame_pairs` -- the order in which `TrackingDataVOD.__getitem__` hands out frame pairs: clip by clip, frame
                        `current + 1` as the FIRST cloud and frame `current` as the second, `curr_idx = current + 1`, the
                        new-sequence flag on the first pair a clip yields, unreadable frames skipped (the reference's bare
                        `except`).  Only the radar records are read: the lidar sweeps, calibration and odometry the reference
                        loads next to them feed its labels and its visualisation, not the model.
* `eval_epoch`       -- per pair: state reset at a new sequence (:69-73), input slicing (:75-79), `net(...)` (:127-130), the
                        carried state detached (:160-165), the per-frame result file (:167-184), metrics if ground truth is
                        supplied (:132-150).  The label pipeline that produces the ground truth (3-D boxes -> moving points,
                        per-point flow) is the dataset's business and stays outside; `gt_fn` is where it plugs in.
* `clips_of_rank`,   -- more than one GPU: the recurrent state and the tracked objects belong to a sequence, so
  `reduce_summary`      whole clips are dealt to the ranks (a sequence never changes rank, no data-path collective) and the per-rank
                        summaries are summed with one small all_reduce at the end.

Differences from the reference, on purpose: the two clouds of a real pair differ in size (242-352 points in the example set);
the reference feeds them as they are, one pair per call.  Here a pair of unequal sizes goes through the variable-size entry
(`Track4D.forward(..., npts1=, npts2=)`, zero padding + point counts: data_io.PaddedBatcher's convention), equal sizes through
the plain one.  Metrics stay on the device until the end (metrics.py); the reference reads them back every frame.
The reference loads the tracking labels of both frames before it calls the model and skips the frame when that fails (`except:
continue`, :96-108), so its loop cannot run on unlabelled data at all; here every pair goes through the model and ground truth only
gates the metrics.
"""
import os

import numpy as np
import torch

from . import data_io, metrics

FLOW_KEYS = ("rne", "50-50 rne", "mov_rne", "stat_rne", "sas", "ras", "epe")      # eval_scene_flow, main_utils.py:342-374
SEG_KEYS = ("acc", "miou", "sen")                                                 # eval_motion_seg, main_utils.py:377-389


def clips_of_rank(clips, rank, world):
    """The clips rank `rank` of `world` evaluates: dealt round-robin in list order, every clip to exactly one rank."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside a world of {world}")
    return list(clips[rank::world])


def reduce_summary(summary, group=None, device=None):
    """Sum what eval_epoch returned over the ranks of `group` (torch.distributed, already initialised): counts and metric sums
    add up, as the reference's single loop would have accumulated them.  One all_reduce of 13 doubles; `device` = where the
    collective's buffer lives (a CUDA device for NCCL; default CPU for gloo).  Without an initialised process group: a copy."""
    import torch.distributed as dist

    vals = [summary["frames"], summary["objects"], summary["examples"]]
    vals += [summary["flow"].get(k, 0.0) for k in FLOW_KEYS] + [summary["seg"].get(k, 0.0) for k in SEG_KEYS]
    t = torch.tensor(vals, dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, group=group)
    v = t.cpu().tolist()
    out = {"frames": int(round(v[0])), "objects": int(round(v[1])), "examples": int(round(v[2])), "flow": {}, "seg": {}}
    if out["examples"]:
        out["flow"] = dict(zip(FLOW_KEYS, v[3:3 + len(FLOW_KEYS)]))
        out["seg"] = dict(zip(SEG_KEYS, v[3 + len(FLOW_KEYS):]))
    return out


def clip_frame_pairs(radar_dir, clips_dir, clips, reader=data_io.read_radar_bin, need_previous=True):
    """Yield (frame_a (Na,7), frame_b (Nb,7), index, clip, is_new_seq) in the reference's order.

    radar_dir: directory of `<frame:05d>.bin` radar records (VodTrackLocations.radar_dir); clips_dir: directory of `<clip>.txt`
    files listing a clip's frame numbers, first and last line = its range (track_vod_3d.py:56-63); clips: clip names in order
    (the reference's train / val / test lists, :33-35).
    need_previous: the reference also loads frame `current - 1` (its lidar sweep, for the visualisation, :75-76,89) inside the
    same try block, so a pair whose previous frame does not exist is dropped -- in practice the pair at frame 0 of the data set.
    True mirrors that (the previous frame's radar record must exist); False yields every readable pair.
    One deliberate difference: past the end of a clip whose last frames are unreadable the reference keeps counting up into the
    next clip's numbers (its `while True` never re-checks the range); here a clip ends at its last line."""
    for clip in clips:
        with open(os.path.join(clips_dir, clip + ".txt")) as f:
            lines = f.read().splitlines()
        first, last = int(lines[0]), int(lines[-1])
        current, new_seq = first, True
        while current + 1 <= last:
            try:
                a = reader(os.path.join(radar_dir, str(current + 1).zfill(5) + ".bin"))
                b = reader(os.path.join(radar_dir, str(current).zfill(5) + ".bin"))
                if need_previous and not os.path.exists(os.path.join(radar_dir, str(current - 1).zfill(5) + ".bin")):
                    raise OSError("no previous frame")
            except (OSError, ValueError):
                current += 1                       # the reference's `except: self.current_frame += 1`
                continue
            current += 1
            yield a, b, current, clip, new_seq
            new_seq = False


def pair_tensors(frame_a, frame_b, device=None, multiple=4):
    """(Na,7), (Nb,7) records -> pc1, pc2 (1,3,N), ft1, ft2 (1,2,N) float32 tensors on `device` and npts1, npts2:
    None for equal sizes, else (1,) int32 counts with N = max(Na, Nb) rounded up to `multiple` and zero padding."""
    a, b = np.asarray(frame_a, np.float32), np.asarray(frame_b, np.float32)
    if a.shape[0] == b.shape[0]:
        t = [torch.from_numpy(x) for x in data_io.frame_pair_inputs(a, b)]
        n1 = n2 = None
    else:
        batcher = data_io.PaddedBatcher(batch=1, pin=False, multiple=multiple)
        batcher.add(0, a, b)
        (_, *t, n1, n2), = list(batcher.flush())
    if device is not None:
        t = [x.to(device, non_blocking=True) for x in t]
    return t[0], t[1], t[2], t[3], n1, n2


def write_result_file(results_dir, clip, index, objects, confs):
    """`<results_dir>/<clip>/<index:05d>.txt`, one line per tracked object in dict order with the confidence of the same
    position (main_utils.py:167-184).  -> the path."""
    d = os.path.join(results_dir, clip)
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, str(index).zfill(5) + ".txt")
    with open(path, "w") as f:
        for i, (obj_id, obj) in enumerate(objects.items()):
            f.write(data_io.format_result_line(obj_id, confs[i], obj))
    return path


@torch.no_grad()
def eval_epoch(net, pairs, results_dir=None, device=None, gt_fn=None):
    """Run `net` (a Track4D in eval mode) over `pairs` (what clip_frame_pairs yields), carrying the recurrent state `h` and the
    tracked objects from frame to frame within a sequence.

    gt_fn(clip, index, pc1) -> None or (gt_flow (1,3,N), gt_cls (1,N)) on pc1's device: frames with ground truth enter the
    scene-flow and motion-segmentation metrics (sums, as the reference accumulates them, plus their count).
    -> {"frames": n, "objects": tracked objects written, "flow": {...}, "seg": {...}, "examples": frames with ground truth}"""
    objects_prev, h = dict(), None
    flow_met, seg_met = dict(), dict()
    n_frames = n_objects = n_gt = 0
    for frame_a, frame_b, index, clip, is_new_seq in pairs:
        if is_new_seq:
            objects_prev, h = dict(), None
        pc1, pc2, ft1, ft2, n1, n2 = pair_tensors(frame_a, frame_b, device)
        kw = {} if n1 is None else {"npts1": n1, "npts2": n2}
        h, pc1_warp, cls, _, _, _, confs, objects, _, _ = net(pc1, pc2, ft1, ft2, h, objects_prev, **kw)
        if n1 is not None:
            pc1 = pc1[..., :int(n1[0])]
        n_frames += 1
        n_objects += len(objects)
        if gt_fn is not None:
            gt = gt_fn(clip, index, pc1)
            if gt is not None:
                gt_flow, gt_cls = gt
                metrics.accumulate(flow_met, metrics.eval_scene_flow(pc1, pc1_warp, gt_flow, cls))
                cls_mask = (cls.reshape(gt_cls.shape) > 0.5).float()
                metrics.accumulate(seg_met, metrics.eval_motion_seg(cls_mask, gt_cls.float()))
                n_gt += 1
        objects_prev = {k: v.clone().detach() for k, v in objects.items()}
        if h is not None:
            h = h.detach()
        if results_dir is not None:
            write_result_file(results_dir, clip, index, objects, confs)
    return {"frames": n_frames, "objects": n_objects, "examples": n_gt,
            "flow": metrics.as_floats(flow_met), "seg": metrics.as_floats(seg_met)}
