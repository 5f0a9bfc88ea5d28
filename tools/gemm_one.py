#!/usr/bin/env python
"""Runs one tensor-core primitive a few times (for ncu captures).
  python tools/gemm_one.py fwd N H W Cin Cout KH KW [reps]      e.g.  fwd 1 1 16384 320 2560 1 1
  python tools/gemm_one.py wgrad N H W Cin Cout KH KW [reps]
  python tools/gemm_one.py dgrad N H W Cin Cout KH KW [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from t2v_b200 import prims  # noqa: E402

kind = sys.argv[1]
N, H, W, Ci, Co, KH, KW = (int(v) for v in sys.argv[2:9])
reps = int(sys.argv[9]) if len(sys.argv) > 9 else 5
with_stats = "stats" in sys.argv      # forward only: also produce the per-frame GroupNorm statistics in the epilogue
with_dbias = "dbias" in sys.argv      # wgrad only: also produce the bias gradient (row sums of dy^T) in the same launch
pads = ((KH - 1) // 2, (KH - 1) // 2, (KW - 1) // 2, (KW - 1) // 2)
dev = "cuda"
x = torch.randn(N, H, W, Ci, device=dev).bfloat16()
w = (torch.randn(Co, KH, KW, Ci, device=dev) * 0.02).bfloat16()
bias = torch.randn(Co, device=dev)
dy = torch.randn(N, H, W, Co, device=dev).bfloat16()
dw = torch.zeros(Co, KH, KW, Ci, device=dev)
db = torch.zeros(Co, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps + 2):
    if i == 2:
        e0.record()
    if kind == "fwd":
        st = prims.stats_alloc(N, Co, dev) if with_stats else None
        prims.conv_fwd(x, w, bias, None, None, 1, pads, stats=st, stats_rows=H * W if with_stats else 0)
    elif kind == "dgrad":
        prims.conv_dgrad(dy, w, (H, W), 1, pads)
    else:
        prims.conv_wgrad(x, dy, dw, 1, pads, dbias=db if with_dbias else None)
e1.record()
torch.cuda.synchronize()
fl = 2.0 * N * H * W * Co * KH * KW * Ci
us = 1e3 * e0.elapsed_time(e1) / reps
print(f"{kind} {sys.argv[2:9]}: {us:.1f} us  {fl / us / 1e6:.1f} TFLOP/s")
