#!/usr/bin/env python
"""Runs GroupNorm forward + backward a few times on one shape (for ncu captures):  python tools/gn_one.py S P C [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from t2v_b200 import prims  # noqa: E402

S, P, C = (int(v) for v in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
x = torch.randn(S, P, C, device="cuda").bfloat16()
dy = torch.randn(S, P, C, device="cuda").bfloat16()
gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
for _ in range(reps):
    y, stat, ab = prims.groupnorm_fwd(x, gamma, beta, 32, 1e-5, 1)
    prims.groupnorm_bwd(dy, x, gamma, stat, ab, 32, 1, None, dg, db)
torch.cuda.synchronize()
print("ok")
