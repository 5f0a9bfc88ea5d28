#!/usr/bin/env python
"""GPU-time micro-benchmark of every distinct primitive call of one finetune step.

Pass 1 (eager): hook prims.* during one cfg-2 step and record each call's argument template (tensor shapes/dtypes,
scalars).  Pass 2: for each distinct template, rebuild random tensors, capture a CUDA graph of R back-to-back calls and
time its replay - i.e. the kernel time the graph-replayed step actually pays, free of host launch latency.
Prints per-template time, count per step, TFLOP/s or GB/s, and the weighted total.  Usage: python tools/shape_bench.py"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402

import bench  # noqa: E402
from op_profile import sig_and_work  # noqa: E402
from t2v_b200 import prims  # noqa: E402
from t2v_b200 import step as S  # noqa: E402


from t2v_b200.profiling import prim_names, record_calls, replay_us  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--top", type=int, default=80)
    ap.add_argument("--only", default="")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    unet = bench.build_unet(dev, args.small)
    step = S.DataParallelStep(unet, S.ddpm_alphas_cumprod(device=dev), passes=1, use_graph=False)
    inputs = [x.to(dev) for x in bench.synthetic_inputs(1, bench.CFG2, 1234)]
    step(*inputs)
    torch.cuda.synchronize()
    calls = record_calls(lambda: step(*inputs), prim_names(), sig_and_work)
    del step, unet
    torch.cuda.empty_cache()

    rows = []
    for key, (cnt, (sig, fl, by)) in calls.items():
        n = key[0]
        if args.only and args.only not in n:
            continue
        try:
            us = replay_us(key, dev, args.reps)
        except Exception as ex:  # noqa: BLE001
            print("skip", n, sig, repr(ex)[:100])
            continue
        rows.append(dict(prim=n, sig=sig, us=us, n=cnt, flops=fl, bytes=by, ms_per_step=us * cnt / 1e3))
    rows.sort(key=lambda r: -r["ms_per_step"])
    tot = sum(r["ms_per_step"] for r in rows)
    fam = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows:
        fam[r["prim"]][0] += r["n"]
        fam[r["prim"]][1] += r["ms_per_step"]
        fam[r["prim"]][2] += r["flops"] * r["n"]
    print(f"sum over the step of graph-replayed primitive times: {tot:.2f} ms")
    for n, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        extra = f" {v[2] / v[1] / 1e9:8.1f} TFLOP/s" if v[2] else ""
        print(f"  {n:22s} n={v[0]:6d} {v[1]:8.2f} ms {100 * v[1] / tot:5.1f}%{extra}")
    print()
    for r in rows[:args.top]:
        eff = f"{r['flops'] / r['us'] / 1e6:7.1f} TF/s" if r["flops"] else f"{r['bytes'] / r['us'] / 1e3:7.0f} GB/s"
        print(f"  {r['ms_per_step']:7.3f} ms n={r['n']:4d} {r['us']:8.1f} us {eff}  {r['prim']} {r['sig']}")
    if args.json:
        with open(args.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
