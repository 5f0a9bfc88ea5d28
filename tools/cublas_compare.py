#!/usr/bin/env python
"""Where is the practical ceiling?  The plain-GEMM (1x1) shapes of the cfg-2 step, in their three forms, timed through this
library's tcgen05 kernel and through cuBLAS (torch.matmul) with the same methodology: a CUDA graph of back-to-back launches
rotating over operand sets whose footprint exceeds L2, CUDA events around the replays.  cuBLAS is the yardstick
MEASURED_PEAKS.json's `bf16_tflops_sustained` was taken with; this table shows what it reaches on the step's OWN shapes
(bias, residual and GroupNorm statistics are not part of the comparison: cuBLAS would need extra passes for them).

  python tools/cublas_compare.py > gpurun_out/cublas_compare.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from t2v_b200 import prims  # noqa: E402

# (rows M, in K, out N, launches per step of the forward form) - the transformer / resnet 1x1 contractions of cfg 2
SHAPES = [
    (16384, 320, 320, 20), (16384, 320, 960, 15), (16384, 320, 2560, 10), (16384, 1280, 320, 10),
    (4096, 640, 640, 20), (4096, 640, 1920, 15), (4096, 640, 5120, 10), (4096, 2560, 640, 10),
    (1024, 1280, 1280, 20), (1024, 1280, 3840, 15), (1024, 1280, 10240, 10), (1024, 5120, 1280, 10),
    (16384, 2560, 2560, 0), (8192, 8192, 8192, 0),        # two textbook shapes (not in the step) as a sanity anchor
]


def time_graph(fns, reps, replays=3):
    for f in fns[:2]:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fns[i % len(fns)]()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (replays * reps)


def main():
    dev = torch.device("cuda", 0)
    print(f"{'form':6s} {'M':>6s} {'K':>6s} {'N':>6s} | {'ours us':>8s} {'TF/s':>7s} | {'cuBLAS us':>9s} {'TF/s':>7s} | ours/cuBLAS time")
    tot = {"ours": 0.0, "cublas": 0.0}
    for M, K, N, cnt in SHAPES:
        per = 2 * (M * K + N * K + M * N) + 4 * N * K
        nset = min(32, max(4, -(-(300 << 20) // per)))
        reps = max(10, nset)
        xs = [torch.randn(M, K, device=dev).bfloat16() for _ in range(nset)]
        ws = [(torch.randn(N, K, device=dev) * 0.02).bfloat16() for _ in range(nset)]
        dys = [torch.randn(M, N, device=dev).bfloat16() for _ in range(nset)]
        dws = [torch.zeros(N, K, device=dev) for _ in range(nset)]
        outs_y = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(nset)]
        outs_x = [torch.empty(M, K, device=dev, dtype=torch.bfloat16) for _ in range(nset)]
        outs_w = [torch.empty(N, K, device=dev, dtype=torch.bfloat16) for _ in range(nset)]
        fl = 2.0 * M * K * N
        forms = {
            "fwd": ([lambda i=i: prims.conv_fwd(xs[i].view(1, 1, M, K), ws[i].view(N, 1, 1, K)) for i in range(nset)],
                    [lambda i=i: torch.mm(xs[i], ws[i].t(), out=outs_y[i]) for i in range(nset)]),
            "dgrad": ([lambda i=i: prims.conv_dgrad(dys[i].view(1, 1, M, N), ws[i].view(N, 1, 1, K), (1, M)) for i in range(nset)],
                      [lambda i=i: torch.mm(dys[i], ws[i], out=outs_x[i]) for i in range(nset)]),
            "wgrad": ([lambda i=i: prims.conv_wgrad(xs[i].view(1, 1, M, K), dys[i].view(1, 1, M, N), dws[i].view(N, 1, 1, K)) for i in range(nset)],
                      [lambda i=i: torch.mm(dys[i].t(), xs[i], out=outs_w[i]) for i in range(nset)]),
        }
        for name, (ours, cub) in forms.items():
            a = time_graph(ours, reps)
            b = time_graph(cub, reps)
            tot["ours"] += a * cnt
            tot["cublas"] += b * cnt
            print(f"{name:6s} {M:6d} {K:6d} {N:6d} | {a:8.1f} {fl / a / 1e6:7.1f} | {b:9.1f} {fl / b / 1e6:7.1f} | {a / b:5.2f}   (x{cnt}/step)")
        del xs, ws, dys, dws, outs_y, outs_x, outs_w
        torch.cuda.empty_cache()
    print(f"\nweighted by launches per step (fwd count used for all three forms): ours {tot['ours'] / 1e3:.2f} ms, cuBLAS {tot['cublas'] / 1e3:.2f} ms")
    print("note: the wgrad of this library accumulates into fp32 (red.add into the gradient arena); cuBLAS writes bf16 here.")


if __name__ == "__main__":
    main()
