#!/usr/bin/env python
"""Sweeps the planner overrides (T2V_FORCE_BN / T2V_FORCE_MH / T2V_FORCE_SPLITS) for one conv/linear GEMM shape and
prints graph-replayed device time per launch for each, next to the planner's own choice.
  python tools/gemm_sweep.py <fwd|dgrad|wgrad> N H W Cin Cout KH KW"""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from t2v_b200 import prims  # noqa: E402


def timed(fn, reps=10):
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


def main():
    kind = sys.argv[1]
    N, H, W, Ci, Co, KH, KW = (int(v) for v in sys.argv[2:9])
    pads = ((KH - 1) // 2, (KH - 1) // 2, (KW - 1) // 2, (KW - 1) // 2)
    dev = "cuda"
    x = torch.randn(N, H, W, Ci, device=dev).bfloat16()
    w = (torch.randn(Co, KH, KW, Ci, device=dev) * 0.02).bfloat16()
    bias = torch.randn(Co, device=dev)
    dy = torch.randn(N, H, W, Co, device=dev).bfloat16()
    dw = torch.zeros(Co, KH, KW, Ci, device=dev)
    fl = 2.0 * N * H * W * Co * KH * KW * Ci

    def run():
        if kind == "fwd":
            prims.conv_fwd(x, w, bias, None, None, 1, pads)
        elif kind == "dgrad":
            prims.conv_dgrad(dy, w, (H, W), 1, pads)
        else:
            prims.conv_wgrad(x, dy, dw, 1, pads)

    for k in ("T2V_FORCE_BN", "T2V_FORCE_MH", "T2V_FORCE_SPLITS"):
        os.environ.pop(k, None)
    base = timed(run)
    print(f"{kind} {sys.argv[2:9]}  planner: {base:.1f} us  {fl / base / 1e6:.1f} TF/s")
    ncols = Ci if kind in ("dgrad", "wgrad") else Co
    bns = sorted({b for b in (64, 80, 96, 112, 128, 160, 192, 208, 224, 240, 256) if b - 16 < ncols} | {min(256, (ncols + 15) // 16 * 16)})
    splits = (0,) if kind != "wgrad" else (1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64)
    rows = []
    for bn, mh, sp in itertools.product(bns, (1, 2), splits):
        os.environ["T2V_FORCE_BN"], os.environ["T2V_FORCE_MH"] = str(bn), str(mh)
        if sp:
            os.environ["T2V_FORCE_SPLITS"] = str(sp)
        try:
            us = timed(run)
        except Exception as ex:  # noqa: BLE001
            torch.cuda.synchronize()
            rows.append((1e9, bn, mh, sp, repr(ex)[:60]))
            continue
        rows.append((us, bn, mh, sp, ""))
    rows.sort()
    if os.environ.get("SWEEP_ALL"):
        for us, bn, mh, sp, err in sorted(rows, key=lambda r: (r[2], r[1], r[3])):
            print(f"CSV,{kind},{','.join(sys.argv[2:9])},{bn},{mh},{sp},{us:.2f}")
    for us, bn, mh, sp, err in rows[:8]:
        print(f"   bn={bn:3d} mh={mh} splits={sp:3d}: {us:7.1f} us  {fl / us / 1e6:7.1f} TF/s {err}")


if __name__ == "__main__":
    main()
