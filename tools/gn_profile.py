#!/usr/bin/env python
"""GroupNorm forward (with producer statistics) + backward on the step's characteristic shapes: graph-replay timing of each
call, for ncu launch lists / full captures.   python tools/gn_profile.py [--time]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from t2v_b200 import prims  # noqa: E402

SHAPES = [(16, 1024, 320, 1), (1, 16384, 320, 16), (1, 256, 1280, 16), (1, 1024, 1280, 16), (16, 16, 1280, 1), (1, 4096, 640, 16)]


def graph_us(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (3 * reps)


for S, P, C, fps in SHAPES:
    x = torch.randn(S, P, C, device="cuda").bfloat16()
    dy = torch.randn(S, P, C, device="cuda").bfloat16()
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    st = prims.channel_stats(x.view(S * fps, P // fps, C))
    y, stat, ab = prims.groupnorm_fwd(x, gamma, beta, 32, 1e-5, 1, [st], fps)
    prims.groupnorm_bwd(dy, x, gamma, stat, ab, 32, 1, None, dg, db)
    torch.cuda.synchronize()
    if "--time" in sys.argv:
        f = graph_us(lambda: prims.groupnorm_fwd(x, gamma, beta, 32, 1e-5, 1, [st], fps))
        f0 = graph_us(lambda: prims.groupnorm_fwd(x, gamma, beta, 32, 1e-5, 1))
        b = graph_us(lambda: prims.groupnorm_bwd(dy, x, gamma, stat, ab, 32, 1, None, dg, db))
        s = graph_us(lambda: prims.silu_bf16(x.view(-1)))
        print(f"({S},{P},{C}) fps={fps}: fwd(stats) {f:6.1f} us  fwd(own sums) {f0:6.1f} us  bwd {b:6.1f} us   [silu elementwise {s:5.1f} us]", flush=True)
print("ok")
