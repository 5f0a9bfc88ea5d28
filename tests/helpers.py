"""Shared test helpers: CPU emulation switch, seeded state dicts, error metrics."""
import contextlib

import torch


@contextlib.contextmanager
def emulated_prims():
    """Monkeypatch t2v_b200.prims with the fp32 torch restatement in oracle/ops_ref.py (tests only, CPU friendly)."""
    from oracle import ops_ref
    from t2v_b200 import prims
    saved = {}
    for name in ops_ref.ALL:
        if hasattr(prims, name):
            saved[name] = getattr(prims, name)
            setattr(prims, name, getattr(ops_ref, name))
    try:
        yield
    finally:
        for name, fn in saved.items():
            setattr(prims, name, fn)


def seeded_state_dict(model, seed=0, conv4_std=0.02):
    """Deterministic weights for parity tests, independent of module declaration order (each tensor is drawn from a
    generator seeded by crc32(key) ^ seed).  Norm scales ~ 1 + 0.1 N, biases ~ 0.05 N, matrices ~ N / sqrt(fan_in);
    the zero-initialised TemporalConvLayer.conv4 is re-drawn N(0, std) so the temporal-conv branch carries signal."""
    import zlib
    sd = {}
    for k, v in model.state_dict().items():
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
        shape = tuple(v.shape)
        if ".conv4.3." in k:
            sd[k] = torch.randn(shape, generator=g) * conv4_std
        elif v.dim() <= 1:
            sd[k] = 1.0 + 0.1 * torch.randn(shape, generator=g) if k.endswith("weight") else 0.05 * torch.randn(shape, generator=g)
        else:
            sd[k] = torch.randn(shape, generator=g) / (v[0].numel() ** 0.5)
    return sd


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30)).item()
