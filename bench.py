#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: finetune frames/sec of the text-to-video-ms-1.7b UNet
(16 frames x 256^2 -> latents 1x4x16x32x32, bf16 compute, full fine-tune: forward + backward of one UNet pass per
step, fp32 gradients, one gradient all-reduce per step when N > 1), data-parallel by clip.

  python bench.py --gpus N --steps K --warmup W            -> one JSON line (rank 0)
  python bench.py --impl reference ...                     -> the reference algorithm on the host CPU cores
                                                              (oracle port; the reference's diffusers path cannot be
                                                              installed here - see DESIGN.md), same metric/unit.
Everything under oracle/ is used only for the cpu_baseline / --impl reference leg.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG2 = dict(model="text-to-video-ms-1.7b UNet3DConditionModel (random init, conv4 re-drawn N(0,0.01))", frames=16, latent_hw=(32, 32),
            text_len=77, text_dim=1024)
FWD_TFLOP_PER_CLIP = 4.887      # algorithmic 2*MAC of conv/linear/attention contractions, cfg 2
PASS_TFLOP_PER_CLIP = 14.66     # forward + backward (dgrad + wgrad), full fine-tune
# of which the spatial self / cross attention products (12 L^2 C per frame and layer fwd+bwd, 12 Lq Lk C for cross attention):
# 0.367 + 0.043 TFLOP.  They leave gemm_tc_kernel when the fused attention kernels are switched on (T2V_FLASH_ATTN=1).
ATTN_TFLOP_PER_CLIP = 0.41


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1443.0), d.get("hbm_gbs", 6569.3), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        mx = max(int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit())
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


def synthetic_inputs(batch, cfg, seed, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    F, (H, W) = cfg["frames"], cfg["latent_hw"]
    lat = torch.randn(batch, 4, F, H, W, generator=g) * 0.18215
    noise = torch.randn(batch, 4, F, H, W, generator=g)
    t = torch.randint(0, 1000, (batch,), generator=g)
    ehs = torch.randn(batch, cfg["text_len"], cfg["text_dim"], generator=g)
    out = [lat, noise, t, ehs]
    if pin:
        out = [x.pin_memory() for x in out]
    return [x.to(device) for x in out] if device != "cpu" else out


def build_unet(device, small=False):
    from t2v_b200.models.unet_3d_condition import UNet3DConditionModel
    kw = dict(block_out_channels=(128, 256, 320, 320), cross_attention_dim=1024) if small else {}
    torch.manual_seed(1234)
    with torch.device(device):
        m = UNet3DConditionModel(**kw)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if ".conv4.3." in n:
                p.normal_(0.0, 0.01)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0  # round 1: dropout-free step (reference eval_train mode, train.py:779-781); stated in config
    return m.train()


def oracle_pass_seconds(sd_cpu, cfg, frames, threads, reps=1):
    """One forward+backward of the reference algorithm (oracle port, fp32) on the host CPU."""
    from oracle import leaves as L
    from oracle import unet3d_ref as R
    torch.set_num_threads(threads)
    c = dict(cfg)
    c["frames"] = frames
    lat, noise, t, ehs = synthetic_inputs(1, c, 99)
    p = {k: v.requires_grad_(True) for k, v in sd_cpu.items()}
    best = None
    for _ in range(reps):
        for v in p.values():
            v.grad = None
        t0 = time.perf_counter()
        loss, _ = R.finetune_loss(p, R.full_config(**c.get("unet_kwargs", {})), lat, noise, t, ehs, L.ddpm_alphas_cumprod())
        loss.backward()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, float(loss)


def gemm_traffic():
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum averaged over the
    launches of one step) from the committed ncu capture profiles/r1_gemm_traffic.json; None if the file is absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def run_reference(args):
    """--impl reference: the reference algorithm (oracle port of train.py:739-834 + UNet wiring) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from t2v_b200.models.unet_3d_condition import UNet3DConditionModel  # parameter shapes only (random init)
    threads = min(os.cpu_count() or 1, 32)  # the oracle's small fp32 ops do not scale past ~32 threads
    cfg = dict(CFG2)
    kw = dict(block_out_channels=(128, 256, 320, 320)) if args.small else {}
    cfg["unet_kwargs"] = kw
    torch.manual_seed(1234)
    m = UNet3DConditionModel(**kw)
    sd = {k: v.detach().clone().contiguous() for k, v in m.state_dict().items()}
    for k in sd:
        if ".conv4.3." in k:
            sd[k].normal_(0.0, 0.01)
    frames = args.ref_frames
    times = []
    for i in range(args.warmup + args.steps):
        dt, _ = oracle_pass_seconds(sd, cfg, frames, threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    fps = frames / (ms / 1e3)
    line = {"impl": "reference", "metric": "finetune frames/sec (one UNet fwd+bwd pass per step)", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: ms-1.7b full finetune 16f 256^2, one pass", "sample": f"{frames}-frame clip at 32x32 latents"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": f"{len(times)} fwd+bwd passes of a {frames}-frame clip (oracle port of the reference algorithm, fp32)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries print (NCCL's version banner, torchrun notes) goes to stderr: stdout carries ONE JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def shutdown(world, step=None):
    """Tear the process group down without hanging: drop the CUDA graph that holds captured NCCL kernels first, and do
    not let communicator destruction or interpreter teardown block the launcher (bounded by a timer)."""
    if world <= 1:
        return
    import gc
    import threading
    sys.stdout.flush()
    sys.stderr.flush()
    threading.Timer(20.0, lambda: os._exit(0)).start()
    try:
        if step is not None:
            step._graph = None
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    finally:
        os._exit(0)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--small", action="store_true", help="debug-size UNet (not a valid bench line)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--ref-frames", type=int, default=4, help="frames of the bounded CPU sample (reference arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from t2v_b200 import native
    from t2v_b200 import step as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    native.lib()  # fail loudly if the CUDA extension is missing

    unet = build_unet(dev, args.small)
    abar = S.ddpm_alphas_cumprod(device=dev)
    step = S.DataParallelStep(unet, abar, passes=1, use_graph=not args.no_graph)
    B = 1
    host = synthetic_inputs(B, CFG2, 1234 + rank, pin=True)
    devin = [x.to(dev) for x in host]
    frames_per_step = world * B * CFG2["frames"]

    # launches per step, counted on an eager step
    eager = S.DataParallelStep(unet, abar, passes=1, use_graph=False, adopt=False)
    eager.arena = step.arena
    eager.sync_gradients = False  # profiling passes below run on their own rank: no collective
    n0 = native.launch_count()
    eager(*devin)
    torch.cuda.synchronize()
    launches_per_step = native.launch_count() - n0

    # ---- dominant-kernel roofline: every tensor-core (implicit-GEMM) launch of the step, timed on the device.  One eager
    # step records each launch's argument template; each distinct template is then replayed as a CUDA graph of
    # back-to-back launches between CUDA events (eager per-launch events would count host launch gaps as kernel time).
    roof = None
    if rank == 0:  # before the step graph is captured (graph-pool memory would distort eager allocation)
        from t2v_b200 import profiling
        calls = profiling.record_calls(lambda: eager(*devin), ["conv_fwd", "conv_dgrad", "conv_wgrad", "bgemm"])
        gemm_ms, n_gemm = 0.0, 0
        for key, (cnt, _) in calls.items():
            gemm_ms += profiling.replay_us(key, dev, reps=5) * cnt / 1e3
            n_gemm += cnt
        torch.cuda.empty_cache()
        peak_tf, peak_hbm, how = peaks()
        from t2v_b200 import ops as _ops
        gemm_tflop = PASS_TFLOP_PER_CLIP - (ATTN_TFLOP_PER_CLIP if _ops._Flash.enabled else 0.0)
        flops = (gemm_tflop if not args.small else float("nan")) * B
        ach = flops / (gemm_ms / 1e3) if gemm_ms > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 implicit-GEMM conv / linear / attention products)",
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": gemm_traffic(),
                "launches": n_gemm, "distinct_shapes": len(calls), "kernel_ms_per_step": gemm_ms, "share_of_step": None,
                "algorithmic_tflop_per_step": flops, "peak_source": how}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg ("value")
    for _ in range(args.warmup):
        loss = step(*devin)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for _ in range(args.steps):
            loss = step(*devin)
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1) / args.steps
    # ---- end-to-end leg: pinned host inputs -> device every step, loss read back every step
    for _ in range(2):
        step(*[x.to(dev, non_blocking=True) for x in host]).item()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        lv = step(*[x.to(dev, non_blocking=True) for x in host]).item()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1) / args.steps
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    if rank != 0:
        shutdown(world, step)
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = min(os.cpu_count() or 1, 32)  # more threads only oversubscribe the small fp32 ops of this model
        sd_cpu = {k: v.detach().float().cpu().contiguous() for k, v in unet.state_dict().items()}
        cfg = dict(CFG2)
        cfg["unet_kwargs"] = dict(block_out_channels=(128, 256, 320, 320)) if args.small else {}
        fr = args.ref_frames
        dt, _ = oracle_pass_seconds(sd_cpu, cfg, fr, threads)
        cpu = {"value": fr / dt, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": f"one fwd+bwd pass of a {fr}-frame clip at 32x32 latents, oracle port of the reference algorithm, fp32, {threads} threads"}

    in_bytes = sum(x.numel() * x.element_size() for x in host)
    line = {
        "metric": "finetune frames/sec (one UNet fwd+bwd pass per step)", "value": frames_per_step / (ms / 1e3), "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "configs[1]: text-to-video-ms-1.7b full finetune, 16 frames 256x256 (latents 1x4x16x32x32 per GPU), bf16",
                   "passes_per_step": 1, "global_batch_clips": world * B, "parallelism": f"dp{world}", "dropout": "off (eval_train)",
                   "l2": "working set (2.8 GB bf16 weights + activations) >> 126 MB L2; no flush needed",
                   "launch_mode": "eager" if args.no_graph else "cuda-graph replay", "small_debug_model": bool(args.small)},
        "e2e": {"value": frames_per_step / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e},
        "gpu_launches": int(launches_per_step * (args.steps)),
        "launches_per_step": int(launches_per_step),
        "clocks": clocks.summary(),
        "loss": lv,
        "roofline": dict(roof, share_of_step=(roof["kernel_ms_per_step"] / ms)) if roof else None,
        "cpu_baseline": cpu,
    }
    emit(line)
    shutdown(world, step)


if __name__ == "__main__":
    main()
