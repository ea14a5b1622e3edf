"""Run the reference's OWN CUDA kernels (oracle/_ref) on the GPU box and dump golden vectors.

TEST INFRASTRUCTURE ONLY.  `gpurun -- python -m oracle.gen_golden_ref_gpu` writes
gpurun_out/ref_gpu_ops.npz; it is then committed as tests/golden/ref_gpu_ops.npz and pins the C
oracle (tests/test_oracle_golden.py::test_c_oracle_matches_reference_kernels_golden) on any host.
Inputs are regenerated from seeds by the tests (ratrack_b200.synthetic.make_batch), only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [(2, 256, 512), (2, 1024, 512), (1, 300, 512), (1, 3000, 512)]   # (B, N, npoint)


def run_case(ref, B, N, S, seed=1234):
    from ratrack_b200 import synthetic

    d = synthetic.make_batch(B, N, seed=seed)
    xyz = torch.from_numpy(np.ascontiguousarray(d["pc1"].transpose(0, 2, 1))).cuda()
    out = {}
    temp = torch.full((B, N), 1e10, device="cuda")
    fps = torch.empty((B, S), dtype=torch.int32, device="cuda")
    ref.furthest_point_sampling_wrapper(B, N, S, xyz, temp, fps)
    out["fps"] = fps.cpu().numpy()
    out["fps_temp"] = temp.cpu().numpy()
    new_xyz = torch.gather(xyz, 1, fps.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    for r, ns in ((2.0, 4), (4.0, 8), (8.0, 16), (16.0, 32)):
        idx = torch.zeros((B, S, ns), dtype=torch.int32, device="cuda")
        ref.ball_query_wrapper(B, N, S, r, ns, new_xyz, xyz, idx)
        out[f"bq_{int(r)}_{ns}"] = idx.cpu().numpy()
    d2 = torch.empty((B, N, 3), device="cuda")
    i3 = torch.empty((B, N, 3), dtype=torch.int32, device="cuda")
    ref.three_nn_wrapper(B, N, S, xyz, new_xyz, d2, i3)
    out["nn_d2"], out["nn_idx"] = d2.cpu().numpy(), i3.cpu().numpy()
    kd = torch.empty((B, S, 16), device="cuda")
    ki = torch.empty((B, S, 16), dtype=torch.int32, device="cuda")
    ref.knn_wrapper(B, S, N, 16, new_xyz, xyz, kd, ki)
    out["knn_d2"], out["knn_idx"] = kd.cpu().numpy(), ki.cpu().numpy()
    # three_interpolate with the reference FP weights (lib/pointnet2_modules.py:142-144)
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn((B, 8, S), generator=g).cuda()
    dist = torch.sqrt(d2)
    rec = 1.0 / (dist + 1e-8)
    w = (rec / rec.sum(2, keepdim=True)).contiguous()
    o = torch.empty((B, 8, N), device="cuda")
    ref.three_interpolate_wrapper(B, 8, S, N, feats, i3, w, o)
    out["interp_w"], out["interp"] = w.cpu().numpy(), o.cpu().numpy()
    torch.cuda.synchronize()
    return out


def main():
    from oracle import ref_gpu

    ref = ref_gpu.load()
    assert ref is not None, "oracle/_ref/pointnet2_cuda.so missing (python -m oracle.build_ref in the dev container)"
    save = {}
    for (B, N, S) in CASES:
        for k, v in run_case(ref, B, N, S).items():
            save[f"b{B}_n{N}_s{S}/{k}"] = v
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_gpu_ops.npz"), **save)
    print("wrote gpurun_out/ref_gpu_ops.npz", len(save), "arrays")


if __name__ == "__main__":
    main()
