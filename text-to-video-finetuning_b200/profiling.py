"""Measurement helpers shared by bench.py and tools/: record the argument template of every primitive call of one
step, then time each distinct template as a CUDA graph of back-to-back launches (device time the graph-replayed step
pays per launch, free of host launch latency; CUDA events on the launching stream)."""
import collections

import torch

from . import prims


def templ(x):
    if torch.is_tensor(x):
        return ("T", tuple(x.shape), str(x.dtype).replace("torch.", ""), tuple(x.stride()))
    if isinstance(x, (tuple, list)):
        return ("L", tuple(templ(v) for v in x))
    return ("V", x)


def build(t, dev):
    kind = t[0]
    if kind == "T":
        shape, dt, stride = t[1], getattr(torch, t[2]), t[3]
        span = 1 + sum((s - 1) * st for s, st in zip(shape, stride)) if all(s > 0 for s in shape) else 0
        if dt == torch.int64:
            base = torch.zeros(span, device=dev, dtype=dt)
        else:
            base = (torch.randn(span, device=dev) * 0.5).to(dt)
        return torch.as_strided(base, shape, stride)
    if kind == "L":
        return tuple(build(v, dev) for v in t[1])
    return t[1]


def record_calls(run, names, work=None):
    """Run `run()` once with prims.<names> hooked.  Returns OrderedDict template -> [count, work(name, args, kwargs)]."""
    saved = {n: getattr(prims, n) for n in names}
    calls = collections.OrderedDict()

    def wrap(n, fn):
        def inner(*a, **k):
            key = (n, templ(a), tuple(sorted((kk, templ(v)) for kk, v in k.items())))
            if key not in calls:
                calls[key] = [0, work(n, a, k) if work else None]
            calls[key][0] += 1
            return fn(*a, **k)
        return inner
    for n in names:
        setattr(prims, n, wrap(n, saved[n]))
    try:
        run()
        torch.cuda.synchronize()
    finally:
        for n in names:
            setattr(prims, n, saved[n])
    return calls


def _nbytes(t):
    if torch.is_tensor(t):
        return t.numel() * t.element_size()
    if isinstance(t, (tuple, list)):
        return sum(_nbytes(v) for v in t)
    return 0


def replay_us(key, dev, reps=10, replays=3, cold=False, batches=1):
    """Average device time (us) of one launch of the recorded call `key`, measured over a replayed CUDA graph of `reps`
    back-to-back launches.  cold=True: the launches rotate over several independent operand sets whose total footprint
    exceeds the 126 MB L2 (at least 4 sets), so no launch finds its operands cached by an earlier one - the situation of
    the real step, where 2.8 GB of weights and the activations of ~3,400 other launches pass through L2 between two uses.
    batches > 1: the measurement (an average over replays x reps launches) is repeated and the MEDIAN batch is returned - a
    sub-millisecond window is otherwise at the mercy of one clock ramp or one interfering process."""
    n, ta, tk = key
    fn = getattr(prims, n)
    a = build(ta, dev)
    k = {kk: build(v, dev) for kk, v in tk}
    sets = [(a, k)]
    if cold:
        per = max(1, _nbytes(a) + sum(_nbytes(v) for v in k.values()))
        want = min(64, max(4, -(-(300 << 20) // per)))
        reps = max(reps, want)
        for _ in range(want - 1):
            sets.append((build(ta, dev), {kk: build(v, dev) for kk, v in tk}))
    for aa, kk in sets[:2]:
        fn(*aa, **kk)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            aa, kk = sets[i % len(sets)]
            fn(*aa, **kk)
    g.replay()
    torch.cuda.synchronize()
    samples = []
    for _ in range(max(1, batches)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(replays):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        samples.append(1e3 * e0.elapsed_time(e1) / (replays * reps))
    samples.sort()
    us = samples[len(samples) // 2]
    del g, a, k, sets
    return us


def prim_names():
    skip = ("out_hw", "groupnorm_ws", "concat_channels", "split_channels")
    return [n for n in dir(prims) if callable(getattr(prims, n)) and not n.startswith("_")
            and getattr(getattr(prims, n), "__module__", "") == prims.__name__ and n not in skip]
