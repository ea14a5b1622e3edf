"""Drive the UNMODIFIED reference dataset class (src/dataset_classes/track_vod_3d.py, TrackingDataVOD.__getitem__) over a
directory of radar `.bin` records and print, as JSON lines, what it hands out: the frame order ratrack_b200.main_utils
.clip_frame_pairs is held to.  TEST INFRASTRUCTURE ONLY, dev container only (needs /root/reference).

    cd <dir with ./clips/<clip>.txt>; python -m oracle.ref_frame_order <radar_dir> <n_items>

What is substituted, in the namespace of the reference module and nowhere else: the View-of-Delft file layer
(`FrameDataLoader`, `FrameTransformMatrix`, `VodTrackLocations`) by stand-ins that read `<radar_dir>/<frame>.bin` for the
radar record and for "a lidar sweep exists" alike, with identity calibration; the control flow under test -- clip order, the
pair (current + 1, current), the index, the new-sequence flag, the bare `except` that skips unreadable frames, the dependence
on frame current - 1 -- is the reference's own code."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    radar_dir, n_items = sys.argv[1], int(sys.argv[2])
    sys.path.insert(0, ROOT)
    from oracle import ref_harness

    ref_harness.install()
    import dataset_classes.track_vod_3d as m

    class Loader:
        def __init__(self, kitti_locations, frame_number):
            path = os.path.join(radar_dir, frame_number + ".bin")
            self.radar_data = np.fromfile(path, dtype=np.float32).reshape(-1, 7)      # a missing file raises: the bare except skips
            self.lidar_data = np.zeros((2, 4), np.float32)

    class Transform:
        def __init__(self, frame_data):
            self.t_lidar_radar = self.t_odom_camera = self.t_camera_radar = np.eye(4)

    m.FrameDataLoader, m.FrameTransformMatrix = Loader, Transform
    m.VodTrackLocations = lambda **kw: kw

    class Args:
        eval = True
        dataset_path = radar_dir

    ds = m.TrackingDataVOD(Args(), None)
    for i in range(n_items):
        pc0, pc1, ft0, ft1, _, idx, clip, _, _, _, _, new_seq = ds[i]
        print(json.dumps({"index": int(idx), "clip": clip, "new_seq": bool(new_seq), "n0": int(pc0.shape[0]), "n1": int(pc1.shape[0]),
                          "sum0": float(np.float64(pc0).sum() + np.float64(ft0[:, :2]).sum()),
                          "sum1": float(np.float64(pc1).sum() + np.float64(ft1[:, :2]).sum())}))


if __name__ == "__main__":
    main()
