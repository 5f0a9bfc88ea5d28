"""Per-kernel GPU parity: every HBM-bound sm_100a kernel vs its plain-PyTorch fp32 restatement (oracle/ops_ref.py)
on the same seeded inputs, through the C ABI.  Floating point: tolerances are bf16 output rounding (2^-8) or tighter."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mods():
    from oracle import ops_ref
    from t2v_b200 import prims
    return prims, ops_ref


def _gen(seed=0):
    return torch.Generator(device=DEV).manual_seed(seed)


def rnd(g, *shape, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(*shape, device=DEV, generator=g) * scale).to(dtype)


def close(a, b, tol, what=""):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)
    assert err < tol, f"{what}: rel-to-max error {err:.3e} >= {tol}"


@pytest.mark.parametrize("S,P,C,G,silu", [(16, 1024, 320, 32, 1), (1, 4096, 640, 32, 1), (2, 64, 2560, 32, 1), (4, 100, 64, 32, 0),
                                          (3, 17, 1280, 32, 0), (1, 16384, 128, 32, 1)])
def test_groupnorm(S, P, C, G, silu):
    prims, ref = _mods()
    g = _gen(1)
    x = rnd(g, S, P, C) * 2 + 0.5
    gamma = 1 + 0.2 * torch.randn(C, device=DEV, generator=g)
    beta = 0.1 * torch.randn(C, device=DEV, generator=g)
    eps = 1e-5
    y, stat, ab = prims.groupnorm_fwd(x, gamma, beta, G, eps, silu)
    y_r, stat_r, ab_r = ref.groupnorm_fwd(x, gamma, beta, G, eps, silu)
    close(y, y_r, 1e-2, "gn y")
    close(stat, stat_r, 1e-3, "gn stat")
    close(ab, ab_r, 1e-3, "gn ab")
    dy = rnd(g, S, P, C)
    add = rnd(g, S, P, C)
    dg, db = torch.ones(C, device=DEV), torch.ones(C, device=DEV)
    dg_r, db_r = dg.clone(), db.clone()
    dx = prims.groupnorm_bwd(dy, x, gamma, stat, ab, G, silu, add, dg, db)
    dx_r = ref.groupnorm_bwd(dy, x, gamma, stat_r, ab_r, G, silu, add, dg_r, db_r)
    close(dx, dx_r, 1.5e-2, "gn dx")
    close(dg, dg_r, 3e-3, "gn dgamma")
    close(db, db_r, 3e-3, "gn dbeta")


@pytest.mark.parametrize("rows,C", [(16384, 320), (4096, 640), (1000, 1280), (77, 512), (5, 64), (300, 2048)])
def test_layernorm(rows, C):
    prims, ref = _mods()
    g = _gen(2)
    x = rnd(g, rows, C) * 1.5 + 0.3
    gamma = 1 + 0.2 * torch.randn(C, device=DEV, generator=g)
    beta = 0.1 * torch.randn(C, device=DEV, generator=g)
    y, stat = prims.layernorm_fwd(x, gamma, beta, 1e-5)
    y_r, stat_r = ref.layernorm_fwd(x, gamma, beta, 1e-5)
    close(y, y_r, 1e-2, "ln y")
    close(stat, stat_r, 1e-4, "ln stat")
    dy, add = rnd(g, rows, C), rnd(g, rows, C)
    dg, db = torch.ones(C, device=DEV), torch.ones(C, device=DEV)
    dg_r, db_r = dg.clone(), db.clone()
    dx = prims.layernorm_bwd(dy, x, gamma, stat, add, dg, db)
    dx_r = ref.layernorm_bwd(dy, x, gamma, stat_r, add, dg_r, db_r)
    close(dx, dx_r, 1.5e-2, "ln dx")
    close(dg, dg_r, 3e-3, "ln dgamma")
    close(db, db_r, 3e-3, "ln dbeta")


def test_geglu_silu_add_scale_cast():
    prims, ref = _mods()
    g = _gen(3)
    proj = rnd(g, 1000, 2560, scale=2.0)
    close(prims.geglu_fwd(proj), ref.geglu_fwd(proj), 1e-2, "geglu")
    dout = rnd(g, 1000, 1280)
    close(prims.geglu_bwd(proj, dout), ref.geglu_bwd(proj, dout), 1e-2, "geglu bwd")
    x = rnd(g, 3, 1280, scale=3.0)
    close(prims.silu_bf16(x), ref.silu_bf16(x), 1e-2, "silu")
    close(prims.silu_bf16_bwd(x, x), ref.silu_bf16_bwd(x, x), 1e-2, "silu bwd")
    xf = torch.randn(3, 1280, device=DEV, generator=g)
    close(prims.silu_f32_to_bf16(xf), ref.silu_f32_to_bf16(xf), 1e-2, "silu f32")
    close(prims.silu_bwd_f32(xf, xf), ref.silu_bwd_f32(xf, xf), 1e-4, "silu bwd f32")
    a, b, c = rnd(g, 40, 64), rnd(g, 40, 64), rnd(g, 40, 64)
    close(prims.add_bf16(a, b, c), ref.add_bf16(a, b, c), 1e-2, "add3")
    close(prims.add_bf16(a, b), ref.add_bf16(a, b), 1e-2, "add2")
    close(prims.scale_bf16(a, 0.37), ref.scale_bf16(a, 0.37), 1e-2, "scale")
    src = torch.randn(1003, device=DEV, generator=g)
    assert torch.equal(prims.cast_f32_bf16(src), src.bfloat16())


@pytest.mark.parametrize("N,H,W,Ho,Wo,C", [(4, 8, 8, 16, 16, 64), (2, 2, 2, 3, 3, 128), (1, 5, 9, 10, 18, 8)])
def test_upsample(N, H, W, Ho, Wo, C):
    prims, ref = _mods()
    g = _gen(4)
    x = rnd(g, N, H, W, C)
    assert torch.equal(prims.upsample_nearest_fwd(x, (Ho, Wo)), ref.upsample_nearest_fwd(x, (Ho, Wo)))
    dy = rnd(g, N, Ho, Wo, C)
    close(prims.upsample_nearest_bwd(dy, (H, W)), ref.upsample_nearest_bwd(dy, (H, W)), 1e-2, "upsample bwd")


def test_concat_split_colsum():
    prims, ref = _mods()
    g = _gen(5)
    a, b = rnd(g, 6, 5, 7, 320), rnd(g, 6, 5, 7, 640)
    cat = prims.concat_channels(a, b)
    assert torch.equal(cat, torch.cat([a, b], -1))
    a2, b2 = prims.split_channels(cat, 320)
    assert torch.equal(a2, a) and torch.equal(b2, b)
    x = rnd(g, 4, 300, 320)
    out, out_r = torch.ones(4, 320, device=DEV), torch.ones(4, 320, device=DEV)
    prims.colsum(x, out, 4, 300, 320)
    ref.colsum(x, out_r, 4, 300, 320)
    close(out, out_r, 1e-3, "colsum")
    acc, acc_r = torch.ones(320, device=DEV), torch.ones(320, device=DEV)
    prims.colsum_f32(out, acc)
    ref.colsum_f32(out_r, acc_r)
    close(acc, acc_r, 1e-3, "colsum_f32")


@pytest.mark.parametrize("rows,n,ld", [(5000, 77, 80), (2048, 1024, 1024), (33, 16, 16)])
def test_softmax(rows, n, ld):
    prims, ref = _mods()
    g = _gen(6)
    s = torch.randn(rows, ld, device=DEV, generator=g) * 4
    p = prims.softmax_fwd(s, n, ld)
    p_r = ref.softmax_fwd(s, n, ld)
    close(p, p_r, 1e-2, "softmax")
    assert (p[:, n:] == 0).all()
    dp = torch.randn(rows, ld, device=DEV, generator=g)
    close(prims.softmax_bwd(p_r, dp, n, 0.125), ref.softmax_bwd(p_r, dp, n, 0.125), 1e-2, "softmax bwd")


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("B,F,HW,heads,D", [(1, 16, 64, 5, 64), (2, 8, 16, 2, 64), (1, 24, 9, 1, 32), (1, 32, 4, 8, 64), (1, 1, 16, 2, 64)])
def test_temporal_attention(B, F, HW, heads, D, fused):
    """fused: q | k | v are column slices of one [rows, 3C] projection (row pitch 3C in, C out)."""
    prims, ref = _mods()
    g = _gen(7)
    C = heads * D
    rows = B * F * HW
    if fused:
        qkv = rnd(g, rows, 3 * C)
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    else:
        q, k, v = (rnd(g, rows, C) for _ in range(3))
    do = rnd(g, rows, C)
    ld_in = 3 * C if fused else C
    addr = (B * HW, HW, F * HW, 1, HW, ld_in, C, heads, F, D)
    o, o_r = torch.zeros_like(do), torch.zeros_like(do)
    close(prims.attn_small_fwd(q, k, v, o, addr), ref.attn_small_fwd(q, k, v, o_r, addr), 1e-2, "attn_small fwd")
    # gradients land in the layout of the inputs (the fused [rows, 3C] buffer when fused)
    got, exp = torch.zeros(rows, 3 * C, device=DEV, dtype=do.dtype), torch.zeros(rows, 3 * C, device=DEV, dtype=do.dtype)
    if fused:
        gaddr = addr
        sl = lambda t: (t[:, :C], t[:, C:2 * C], t[:, 2 * C:])
    else:
        gaddr = addr
        sl = lambda t: tuple(t.view(3, rows, C)[i] for i in range(3))
    prims.attn_small_bwd(q, k, v, do, *sl(got), gaddr)
    ref.attn_small_bwd(q, k, v, do, *sl(exp), gaddr)
    for name, a, b in zip("qkv", sl(got), sl(exp)):
        close(a, b, 1.5e-2, f"attn_small d{name}")


@pytest.mark.parametrize("cross", [False, True])
def test_fused_projection_attention_matches_unfused(cross):
    """ops.attention_fused on [.., 3C] / [.., 2C] projections == ops.attention on separate q, k, v (values and gradients)."""
    from t2v_b200 import ops
    g = _gen(21)
    Nb, Lq, Lk, heads, D = (2, 2304, 77, 5, 64) if cross else (3, 256, 256, 5, 64)
    C = heads * D
    q, k, v = rnd(g, Nb, Lq, C), rnd(g, Nb, Lk, C), rnd(g, Nb, Lk, C)
    do = rnd(g, Nb, Lq, C)
    qs, ks, vs = (t.clone().requires_grad_(True) for t in (q, k, v))
    ops.attention(qs, ks, vs, heads).backward(do)
    if cross:
        a, b = q.clone().requires_grad_(True), torch.cat([k, v], -1).requires_grad_(True)
        out = ops.attention_fused(a, b, heads)
        out.backward(do)
        gq, gk, gv = a.grad, b.grad[..., :C], b.grad[..., C:]
    else:
        a = torch.cat([q, k, v], -1).requires_grad_(True)
        out = ops.attention_fused(a, None, heads)
        out.backward(do)
        gq, gk, gv = a.grad[..., :C], a.grad[..., C:2 * C], a.grad[..., 2 * C:]
    close(out, ops.attention(q, k, v, heads), 1e-2, "fused attention fwd")
    for name, x, y in (("q", gq, qs.grad), ("k", gk, ks.grad), ("v", gv, vs.grad)):
        close(x, y, 1.5e-2, f"fused attention d{name}")


@pytest.mark.parametrize("p,scale", [(0.1, 1.0), (0.5, 0.25), (0.0, 1.0)])
def test_dropout_scale_add_mask_is_bit_exact(p, scale):
    """The dropout mask is a pure function of (seed, element index): the kernel and its torch restatement must keep
    exactly the same elements (integer hash => bit-exact mask), values agree to bf16 rounding."""
    prims, ref = _mods()
    g = _gen(9)
    x, base = rnd(g, 333, 640), rnd(g, 333, 640)
    for seed in (1, 0x9E3779B97F4A7C15 & 0x7FFFFFFFFFFFFFFF):
        y, y_r = prims.dropout_scale_add(x, None, p, scale, seed), ref.dropout_scale_add(x, None, p, scale, seed)
        assert torch.equal(y == 0, y_r == 0), "dropout masks differ"
        close(y, y_r, 1e-2, "dropout values")
        kept = (y != 0).float().mean().item()
        assert abs(kept - (1.0 - p)) < 0.01, kept
        close(prims.dropout_scale_add(x, base, p, scale, seed), ref.dropout_scale_add(x, base, p, scale, seed), 1e-2, "dropout + base")


def test_latent_boundary_and_loss():
    prims, ref = _mods()
    g = _gen(8)
    B, C, F, H, W = 2, 4, 3, 8, 12
    x0 = torch.randn(B, C, F, H, W, device=DEV, generator=g)
    noise = torch.randn(B, C, F, H, W, device=DEV, generator=g)
    abar = torch.linspace(0.999, 0.01, 1000, device=DEV)
    t = torch.tensor([7, 912], device=DEV)
    close(prims.latents_to_nhwc8(x0), ref.latents_to_nhwc8(x0), 1e-2, "to_nhwc8")
    xn = prims.latents_to_nhwc8(x0, noise, abar, t)
    close(xn, ref.latents_to_nhwc8(x0, noise, abar, t), 1e-2, "add_noise")
    assert (xn[..., C:] == 0).all()
    assert torch.equal(prims.nhwc8_to_latents(xn, B, C, F), ref.nhwc8_to_latents(xn, B, C, F))
    loss, loss_r = prims.mse_loss_fwd(xn, noise), ref.mse_loss_fwd(xn, noise)
    assert abs(loss.item() - loss_r.item()) < 1e-5 * abs(loss_r.item()) + 1e-7
    gout = torch.tensor(1.7, device=DEV)
    close(prims.mse_loss_bwd(xn, noise, gout), ref.mse_loss_bwd(xn, noise, gout), 1e-2, "mse bwd")
    tt = torch.tensor([0, 1, 500, 999], device=DEV)
    close(prims.timestep_embedding(tt, 320), ref.timestep_embedding(tt, 320), 1e-2, "timestep emb")


@pytest.mark.parametrize("Nb,Lq,Lk,heads,D", [(4, 256, 256, 2, 64), (2, 1024, 77, 5, 64), (3, 64, 64, 1, 512), (2, 100, 36, 3, 64),
                                            (1, 4096, 77, 5, 64)])  # last: cross-attention split-K gradient path
def test_attention_composite(Nb, Lq, Lk, heads, D):
    """ops.attention (4 tcgen05 batched GEMMs + softmax) vs torch reference, forward and backward."""
    from t2v_b200 import ops
    g = _gen(9)
    C = heads * D
    q = rnd(g, Nb, Lq, C).requires_grad_(True)
    k = rnd(g, Nb, Lk, C).requires_grad_(True)
    v = rnd(g, Nb, Lk, C).requires_grad_(True)
    do = rnd(g, Nb, Lq, C)
    o = ops.attention(q, k, v, heads)
    o.backward(do)
    qf, kf, vf = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    Q = qf.view(Nb, Lq, heads, D).transpose(1, 2)
    K = kf.view(Nb, Lk, heads, D).transpose(1, 2)
    V = vf.view(Nb, Lk, heads, D).transpose(1, 2)
    O = (torch.softmax(Q @ K.transpose(-1, -2) * D ** -0.5, -1) @ V).transpose(1, 2).reshape(Nb, Lq, C)
    O.backward(do.float())
    close(o, O, 1.5e-2, "attention o")
    close(q.grad, qf.grad, 2e-2, "attention dq")
    close(k.grad, kf.grad, 2e-2, "attention dk")
    close(v.grad, vf.grad, 2e-2, "attention dv")


# ------------------------------------------------------------------------------------------------ GroupNorm statistics from GEMM epilogues
@pytest.mark.parametrize("case", [
    # (x shape N,H,W,Ci), Co, (KH,KW), stride, pads, stats_rows, residual, note
    ((16, 32, 32, 64), 320, (3, 3), 1, (1, 1, 1, 1), 1024, True, "3x3 conv per frame (32-row segments)"),
    ((16, 4, 4, 128), 256, (3, 3), 1, (1, 1, 1, 1), 16, False, "4x4 maps: two frames per warp (16-row segments) or split-K finish"),
    ((16, 8, 8, 256), 256, (3, 3), 1, (1, 1, 1, 1), 64, True, "8x8 maps, split-K with statistics in the finishing pass"),
    ((1, 1, 4096, 320), 320, (1, 1), 1, (0, 0, 0, 0), 256, True, "linear (proj_out): frame = row // tokens per frame"),
    ((1, 1, 256, 1280), 1280, (1, 1), 1, (0, 0, 0, 0), 16, True, "linear at the 4x4 level"),
    ((2, 8, 256, 64), 128, (3, 1), 1, (1, 1, 0, 0), 256, True, "temporal conv: [B, F, HW, C], frame = (b, f)"),
    ((1, 16, 16, 640), 640, (3, 1), 1, (1, 1, 0, 0), 16, False, "temporal conv at the 4x4 level"),
    ((16, 32, 32, 64), 128, (3, 3), 2, (1, 1, 1, 1), 256, False, "stride-2 downsample conv"),
    ((4, 12, 20, 64), 96, (3, 3), 1, (1, 1, 1, 1), 240, False, "ragged map (40x72-like): falls back to the statistics pass"),
    ((16, 8, 8, 64), 128, (3, 3), 1, (1, 1, 1, 1), 4 * 64, True, "per-clip statistics from a per-frame conv (4 frames per sample)"),
    ((16, 4, 4, 128), 256, (3, 3), 1, (1, 1, 1, 1), 16 * 16, True, "per-clip statistics at the 4x4 level (one sample)"),
    ((16, 32, 32, 8), 320, (3, 3), 1, (1, 1, 1, 1), 16 * 1024, False, "conv_in -> transformer_in (per clip, 8 input channels)"),
    ((1, 1, 16384, 320), 320, (1, 1), 1, (0, 0, 0, 0), 16384, True, "linear, one sample over all rows"),
    ((1, 16, 1024, 64), 128, (3, 1), 1, (1, 1, 0, 0), 4 * 1024, True, "temporal conv, slots of 4 frames (clip_stats_rows)"),
    ((2, 8, 64, 64), 128, (3, 1), 1, (1, 1, 0, 0), 8 * 64, False, "temporal conv, one slot per clip"),
])
def test_conv_epilogue_statistics(case):
    """T2VEpilogue.stats: the per-(frame, channel) sums a GEMM epilogue (or its split-K finishing pass, or the fallback
    pass) accumulates must equal the sums of the tensor it wrote."""
    prims, ref = _mods()
    (N, H, W, Ci), Co, (KH, KW), stride, pads, srows, with_res, _ = case
    g = _gen(40)
    x = rnd(g, N, H, W, Ci)
    w = rnd(g, Co, KH, KW, Ci, scale=(KH * KW * Ci) ** -0.5)
    bias = torch.randn(Co, device=DEV, generator=g) * 0.5 + 1.0     # non-zero mean: the sums carry a large common mode
    Ho, Wo = prims.out_hw(H, W, KH, KW, stride, pads)
    res = rnd(g, N, Ho, Wo, Co) if with_res else None
    frames = N * Ho * Wo // srows
    stats = prims.stats_alloc(frames, Co, x.device)
    y = prims.conv_fwd(x, w, bias, None, res, stride, pads, stats=stats, stats_rows=srows)
    y_plain = prims.conv_fwd(x, w, bias, None, res, stride, pads)
    assert torch.equal(y, y_plain), "the statistics must not change the output"
    yf = y.float().view(frames, srows, Co)
    want = torch.stack([yf.sum(1), (yf * yf).sum(1)], dim=-1)
    # the epilogue sums the fp32 values before bf16 rounding: agreement is at bf16 rounding level of the SUM, not the element
    close(stats[..., 0], want[..., 0], 5e-3, "sum")
    close(stats[..., 1], want[..., 1], 5e-3, "sum of squares")
    # and a GroupNorm fed with them equals a GroupNorm that computes its own sums
    gamma = 1 + 0.2 * torch.randn(Co, device=DEV, generator=g)
    beta = 0.1 * torch.randn(Co, device=DEV, generator=g)
    for samples in {frames, N if frames % N == 0 else frames}:
        x3 = y.view(samples, -1, Co)
        y1, st1, ab1 = prims.groupnorm_fwd(x3, gamma, beta, 32, 1e-5, 1, [stats], frames // samples)
        y0, st0, ab0 = prims.groupnorm_fwd(x3, gamma, beta, 32, 1e-5, 1)
        close(st1, st0, 2e-3, "gn stat from epilogue sums")
        close(y1, y0, 1e-2, "gn y from epilogue sums")


def test_groupnorm_two_statistics_sources_and_channel_stats():
    """Channel concatenation (up blocks): the consumer GroupNorm takes the sums of the two halves from two buffers."""
    prims, ref = _mods()
    g = _gen(41)
    S, P, Ca, Cb = 4, 256, 320, 640
    a, b = rnd(g, S, P, Ca) + 0.3, rnd(g, S, P, Cb) * 2
    x = torch.cat([a, b], dim=-1).contiguous()
    sa, sb = prims.channel_stats(a), prims.channel_stats(b)
    sa_r = ref.channel_stats(a)
    close(sa, sa_r, 1e-4, "channel_stats")
    gamma = 1 + 0.2 * torch.randn(Ca + Cb, device=DEV, generator=g)
    beta = 0.1 * torch.randn(Ca + Cb, device=DEV, generator=g)
    y1, st1, _ = prims.groupnorm_fwd(x, gamma, beta, 32, 1e-5, 1, [sa, sb], 1)
    y0, st0, _ = ref.groupnorm_fwd(x, gamma, beta, 32, 1e-5, 1)
    close(st1, st0, 1e-3, "gn stat (two sources)")
    close(y1, y0, 1e-2, "gn y (two sources)")
    # per-clip norm over 2 frames per sample
    y2, st2, _ = prims.groupnorm_fwd(x.view(2, 2 * P, Ca + Cb), gamma, beta, 32, 1e-5, 0, [sa, sb], 2)
    y2r, st2r, _ = ref.groupnorm_fwd(x.view(2, 2 * P, Ca + Cb), gamma, beta, 32, 1e-5, 0)
    close(st2, st2r, 1e-3, "gn stat (two sources, per clip)")
    close(y2, y2r, 1e-2, "gn y (two sources, per clip)")


@pytest.mark.parametrize("mean,std", [(8.0, 1.0), (30.0, 0.5), (-3.0, 4.0)])
def test_groupnorm_large_common_mode(mean, std):
    """Round-1 verdict (weak #5): the variance is E[x^2] - E[x]^2 from fp32 partial sums combined in fp64.  With |mean| >> std
    the cancellation costs relative accuracy ~ (mean/std)^2 * 2^-24; the rstd must still agree with the two-pass fp32 oracle
    to 2e-3 at mean/std = 60 (3600x amplification of ~1e-7)."""
    prims, ref = _mods()
    g = _gen(42)
    S, P, C = 2, 4096, 320
    x = (torch.randn(S, P, C, device=DEV, generator=g) * std + mean).bfloat16()
    gamma, beta = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    y, stat, _ = prims.groupnorm_fwd(x, gamma, beta, 32, 1e-5, 0)
    y_r, stat_r, _ = ref.groupnorm_fwd(x, gamma, beta, 32, 1e-5, 0)
    close(stat[..., 1], stat_r[..., 1], 2e-3, "rstd under a large common mode")
    close(y, y_r, 2e-2, "y under a large common mode")
