"""Build the UNMODIFIED reference extension `pointnet2_cuda` into oracle/_ref/.

TEST INFRASTRUCTURE ONLY -- nothing under ratrack_b200/ may import this.

The sources are compiled where they lie under /root/reference/src/lib/src
(pointnet2_api.cpp + 4 wrapper .cpp + 4 kernel .cu, the list in
/root/reference/src/lib/setup.py:7-18, nvcc -O2 as in setup.py:19-20); nothing
is copied into this repository.  The result `oracle/_ref/pointnet2_cuda.so` is
git-ignored but travels to the GPU box with the gpurun snapshot, where it is the
GPU-side parity pin for the oracle (tests/golden/ref_gpu_*.npz are its outputs,
produced by oracle/gen_golden_ref_gpu.py) and the "reference kernels on the same
B200" line of the per-op benchmark.  It cannot run in the dev container (no GPU).
"""
import os
import sys

REF = "/root/reference/src/lib/src"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

SOURCES = [
    "pointnet2_api.cpp",
    "ball_query.cpp", "ball_query_gpu.cu",
    "group_points.cpp", "group_points_gpu.cu",
    "interpolate.cpp", "interpolate_gpu.cu",
    "sampling.cpp", "sampling_gpu.cu",
]


def build(verbose: bool = False) -> str:
    so = os.path.join(OUT, "pointnet2_cuda.so")
    if not os.path.isdir(REF):
        if os.path.exists(so):
            return so
        raise FileNotFoundError("reference sources not present and no prebuilt oracle/_ref")
    srcs = [os.path.join(REF, s) for s in SOURCES]
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load
    load(name="pointnet2_cuda", sources=srcs, extra_cflags=["-g", "-w"],
         extra_cuda_cflags=["-O2", "-w"], build_directory=OUT, verbose=verbose,
         is_python_module=False)
    assert os.path.exists(so), so
    return so


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
