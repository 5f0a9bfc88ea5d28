#!/usr/bin/env python
"""Per-primitive profile of one finetune step: wraps every function in prims.py with CUDA events (eager mode), runs the
cfg-2 step a few times and prints, per (primitive, shape signature): launches, total ms, and the effective TFLOP/s
(tensor-core family) or GB/s (HBM-bound family).  Usage: python tools/op_profile.py [--small] [--top 60] [--json out]"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from t2v_b200 import prims  # noqa: E402
from t2v_b200 import step as S  # noqa: E402


def conv_flops(x, w, stride, pads):
    N, H, W, Ci = x.shape
    Co, KH, KW, _ = w.shape
    Ho, Wo = prims.out_hw(H, W, KH, KW, stride, pads)
    return 2.0 * N * Ho * Wo * Co * KH * KW * Ci


def sig_and_work(name, a, k):
    """Returns (signature string, flops, bytes)."""
    if name == "conv_fwd":
        x, w = a[0], a[1]
        st, pads = k.get("stride", a[5] if len(a) > 5 else 1), k.get("pads", a[6] if len(a) > 6 else (0, 0, 0, 0))
        return f"x{tuple(x.shape)} w{tuple(w.shape)} s{st}", conv_flops(x, w, st, pads), 0
    if name == "conv_dgrad":
        dy, w, in_hw = a[0], a[1], a[2]
        st = a[3] if len(a) > 3 else k.get("stride", 1)
        N = dy.shape[0]
        fl = 2.0 * dy.numel() // dy.shape[-1] * w.shape[0] * w.shape[1] * w.shape[2] * w.shape[3]
        return f"dy{tuple(dy.shape)} w{tuple(w.shape)} s{st}", fl, 0
    if name == "conv_wgrad":
        x, dy, dw = a[0], a[1], a[2]
        fl = 2.0 * (dy.numel() // dy.shape[-1]) * dw.numel()
        return f"x{tuple(x.shape)} dw{tuple(dw.shape)}", fl, 0
    if name == "bgemm":
        M, N, K, Z1, Z2 = a[6:11]
        return f"M{M} N{N} K{K} Z{Z1}x{Z2} a{a[1][0]}b{a[3][0]} mode{k.get('out_mode', a[12] if len(a) > 12 else 0)}", 2.0 * M * N * K * Z1 * Z2, 0
    tens = [t for t in list(a) + list(k.values()) if torch.is_tensor(t)]
    byt = sum(t.numel() * t.element_size() for t in tens)
    shp = tuple(tens[0].shape) if tens else ()
    return f"{shp}", 0, byt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--top", type=int, default=70)
    ap.add_argument("--json", default=None)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--double", action="store_true", help="time the second of two back-to-back runs of each primitive")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    unet = bench.build_unet(dev, args.small)
    abar = S.ddpm_alphas_cumprod(device=dev)
    step = S.DataParallelStep(unet, abar, passes=1, use_graph=False)
    inputs = [x.to(dev) for x in bench.synthetic_inputs(1, bench.CFG2, 1234)]
    for _ in range(2):
        step(*inputs)
    torch.cuda.synchronize()
    names = [n for n in dir(prims) if callable(getattr(prims, n)) and not n.startswith("_") and getattr(getattr(prims, n), "__module__", "") == prims.__name__
             and n not in ("out_hw",)]
    saved = {n: getattr(prims, n) for n in names}
    evs = []

    def wrap(n, fn):
        def inner(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if args.double:   # run the primitive twice and time the second run: its launch is queued behind the first
                fn(*a, **k)   # kernel, so the events bracket pure GPU time (not host launch latency) for kernels > ~20 us
            s.record()
            r = fn(*a, **k)
            e.record()
            if n not in ("concat_channels", "split_channels", "groupnorm_ws"):
                evs.append((n, sig_and_work(n, a, k), s, e))
            return r
        return inner
    for n in names:
        setattr(prims, n, wrap(n, saved[n]))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        step(*inputs)
    e1.record()
    torch.cuda.synchronize()
    for n in names:
        setattr(prims, n, saved[n])
    agg = collections.OrderedDict()
    for n, (sig, fl, by), s, e in evs:
        key = (n, sig)
        d = agg.setdefault(key, [0, 0.0, 0.0, 0.0])
        d[0] += 1
        d[1] += s.elapsed_time(e)
        d[2] += fl
        d[3] += by
    reps = args.reps
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    tot = sum(v[1] for v in agg.values()) / reps
    print(f"step wall (eager, instrumented): {e0.elapsed_time(e1) / reps:.2f} ms; sum of primitive times: {tot:.2f} ms")
    fam = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for (n, sig), v in agg.items():
        fam[n][0] += v[0] / reps
        fam[n][1] += v[1] / reps
        fam[n][2] += v[2] / reps
    print("\nby primitive:")
    for n, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        extra = f" {v[2] / v[1] / 1e9:8.1f} TFLOP/s" if v[2] else ""
        print(f"  {n:22s} n={v[0]:7.0f}  {v[1]:8.2f} ms {100 * v[1] / tot:5.1f}%{extra}")
    print("\ntop signatures:")
    out = []
    for (n, sig), v in rows[:args.top]:
        ms = v[1] / reps
        eff = f"{v[2] / v[1] / 1e9:7.1f} TF/s" if v[2] else f"{v[3] / v[1] / 1e6:7.0f} GB/s"
        print(f"  {ms:7.3f} ms n={v[0] / reps:5.0f} avg={1e3 * v[1] / v[0]:7.1f} us {eff}  {n} {sig}")
        out.append(dict(prim=n, sig=sig, ms=ms, n=v[0] / reps, flops=v[2] / reps, bytes=v[3] / reps))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
