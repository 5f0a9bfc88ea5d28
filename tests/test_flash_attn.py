"""Fused attention (csrc/flash_attn.cu) behind ops._Flash.

CPU (always): the host logic of the fused path - dispatch, saved tensors, column-sliced fused QKV / K|V operands and
gradient layouts - over the emulated primitives must reproduce the unfused path exactly.
GPU: the kernels through the C ABI vs the fp32 restatement, and the end-to-end UNet parity of tests/test_unet_gpu.py with
the fused path (the default since round 2) switched on explicitly."""
import pytest
import torch

from helpers import emulated_prims, rel_l2
from oracle import ops_ref



def _attend(flash, fused, cross, q, k, v, do, heads):
    from t2v_b200 import ops
    old = ops._Flash.enabled
    ops._Flash.enabled = flash
    try:
        C = q.shape[-1]
        if fused and not cross:
            a = torch.cat([q, k, v], -1).requires_grad_(True)
            out = ops.attention_fused(a, None, heads)
            out.backward(do)
            return out.detach(), a.grad[..., :C], a.grad[..., C:2 * C], a.grad[..., 2 * C:]
        if fused:
            a, b = q.clone().requires_grad_(True), torch.cat([k, v], -1).requires_grad_(True)
            out = ops.attention_fused(a, b, heads)
            out.backward(do)
            return out.detach(), a.grad, b.grad[..., :C], b.grad[..., C:]
        qs, ks, vs = (t.clone().requires_grad_(True) for t in (q, k, v))
        out = ops.attention(qs, ks, vs, heads)
        out.backward(do)
        return out.detach(), qs.grad, ks.grad, vs.grad
    finally:
        ops._Flash.enabled = old


@pytest.mark.parametrize("fused,cross", [(False, False), (True, False), (True, True), (False, True)])
def test_fused_path_host_logic_matches_unfused_cpu(fused, cross):
    g = torch.Generator().manual_seed(4)
    Nb, Lq, Lk, heads = (2, 40, 7, 2) if cross else (2, 24, 24, 2)
    C = heads * 64
    q, k, v, do = (torch.randn(Nb, L, C, generator=g) for L in (Lq, Lk, Lk, Lq))
    old = ops_ref.BF
    ops_ref.BF = torch.float32
    try:
        with emulated_prims():
            ref = _attend(False, False, cross, q, k, v, do, heads)
            got = _attend(True, fused, cross, q, k, v, do, heads)
    finally:
        ops_ref.BF = old
    for name, a, b in zip(("o", "dq", "dk", "dv"), got, ref):
        assert rel_l2(a, b) < 1e-5, name


@pytest.mark.gpu
@pytest.mark.parametrize("Nb,Lq,Lk,heads,fused", [(3, 256, 256, 5, False), (2, 1024, 1024, 5, True), (2, 200, 200, 2, False),
                                                   (1, 64, 64, 1, False), (2, 2304, 77, 5, True), (1, 16384, 77, 5, True)])
def test_kernels_match_restatement_gpu(Nb, Lq, Lk, heads, fused):
    from t2v_b200 import prims
    g = torch.Generator(device="cuda").manual_seed(0)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g).bfloat16()  # noqa: E731
    C = heads * 64
    if fused and Lq == Lk:
        qkv = rnd(Nb, Lq, 3 * C)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        dqkv = torch.zeros_like(qkv)
        dq, dk, dv = dqkv[..., :C], dqkv[..., C:2 * C], dqkv[..., 2 * C:]
    elif fused:
        q, kv = rnd(Nb, Lq, C), rnd(Nb, Lk, 2 * C)
        k, v = kv[..., :C], kv[..., C:]
        dq, dkv = torch.zeros_like(q), torch.zeros_like(kv)
        dk, dv = dkv[..., :C], dkv[..., C:]
    else:
        q, k, v = rnd(Nb, Lq, C), rnd(Nb, Lk, C), rnd(Nb, Lk, C)
        dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    do = rnd(Nb, Lq, C)
    o, lse = prims.flash_attn_fwd(q, k, v, heads)
    prims.flash_attn_bwd(q, k, v, o, do, lse, heads, dq, dk, dv)
    o_r, lse_r = ops_ref.flash_attn_fwd(q, k, v, heads)
    gq, gk, gv = torch.zeros_like(dq, dtype=torch.float32), torch.zeros_like(dk, dtype=torch.float32), torch.zeros_like(dv, dtype=torch.float32)
    ops_ref.flash_attn_bwd(q, k, v, o_r, do, lse_r, heads, gq, gk, gv)
    close = lambda a, b, tol, what: None if ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item() < tol \
        else pytest.fail(f"{what}: {((a.float() - b.float()).abs().max() / b.float().abs().max()).item():.3e}")  # noqa: E731
    close(o, o_r, 1e-2, "o")
    close(lse, lse_r, 1e-4, "lse")
    close(dq, gq, 1.5e-2, "dq")
    close(dk, gk, 1.5e-2, "dk")
    close(dv, gv, 1.5e-2, "dv")


@pytest.mark.gpu
def test_unet_parity_with_fused_attention_gpu():
    import test_unet_gpu as U
    from t2v_b200 import ops
    old = ops._Flash.enabled
    ops._Flash.enabled = True
    try:
        U._check(*U._case(U.SMALL, 2, 4, (16, 16)))
        U._check(*U._case(U.MEDIUM, 1, 16, (32, 32)))
    finally:
        ops._Flash.enabled = old
