"""One forward / dgrad / wgrad of the tensor-core dense kernels at a training shape (for ncu captures).
    python tools/run_dense.py [rows] [k] [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ratrack_b200 import _cabi  # noqa: E402
from ratrack_b200.lib import dense_tc  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
k = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = int(sys.argv[3]) if len(sys.argv) > 3 else 256
x = torch.randn(rows, k, device="cuda")
w = torch.randn(n, k, device="cuda") / k ** 0.5
dy = torch.randn(rows, n, device="cuda") * 1e-4
dw = torch.empty(n, k, device="cuda")
for _ in range(2):
    y = dense_tc.forward_raw(x, k, w, k, 1, k, n)
    amax = dense_tc.absmax(dy)
    dx = dense_tc.forward_raw(dy, n, w, 1, k, n, k, None, amax)
    _cabi.call("rt_dense_tc_wgrad", rows, n, k, dy.data_ptr(), n, x.data_ptr(), k, amax.data_ptr(), None, dw.data_ptr(),
               torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok", float(y.abs().max()), float(dx.abs().max()), float(dw.abs().max()))
