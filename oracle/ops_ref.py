"""ORACLE (test infrastructure): plain-PyTorch fp32 restatement of every primitive in
text-to-video-finetuning_b200/prims.py, with identical signatures, tensor layouts and rounding points (bf16 storage
between ops, fp32 math inside).  Two uses, both in tests only:
  * per-kernel GPU parity tests compare each sm_100a kernel against the function of the same name here;
  * CPU tests monkeypatch `prims` with this module to check the host-side wiring and the hand-written backward
    composition in ops.py / layers.py against the model oracle, without a GPU.
The product never imports this file (no CPU fallback exists in the product path).
"""
import math

import torch
import torch.nn.functional as F

BF = torch.bfloat16


def out_hw(H, W, KH, KW, stride, pads):
    return (H + pads[0] + pads[1] - KH) // stride + 1, (W + pads[2] + pads[3] - KW) // stride + 1


def _conv_f32(x, w, stride, pads):
    xf = F.pad(x.float().permute(0, 3, 1, 2), (pads[2], pads[3], pads[0], pads[1]))
    return F.conv2d(xf, w.float().permute(0, 3, 1, 2), stride=stride).permute(0, 2, 3, 1)


def stats_alloc(frames, C, device):
    return torch.zeros((frames, C, 2), dtype=torch.float32, device=device)


def zeros_f32(shape, device):
    return torch.zeros(shape, dtype=torch.float32, device=device)


def channel_stats(x):
    xf = x.float()
    return torch.stack([xf.sum(dim=1), (xf * xf).sum(dim=1)], dim=-1).contiguous()


def conv_fwd(x, w, bias=None, rowbias=None, residual=None, stride=1, pads=(0, 0, 0, 0), alpha=1.0, out_fp32=False, rowbias_div=1,
             stats=None, stats_rows=0):
    y = alpha * _conv_f32(x, w, stride, pads)
    if bias is not None:
        y = y + bias
    if rowbias is not None:
        idx = torch.arange(x.shape[0], device=x.device) // rowbias_div
        y = y + rowbias[idx][:, None, None, :]
    if residual is not None:
        y = y + residual.float()
    if stats is not None:   # epilogue statistics: per-(frame, channel) sums of the fp32 values, rows flattened in [N][Ho][Wo] order
        yf = y.reshape(-1, stats_rows, y.shape[-1])
        stats += torch.stack([yf.sum(dim=1), (yf * yf).sum(dim=1)], dim=-1)
    return y.contiguous() if out_fp32 else y.to(BF).contiguous()


@torch.enable_grad()
def conv_dgrad(dy, w, in_hw, stride=1, pads=(0, 0, 0, 0), residual=None):
    N = dy.shape[0]
    Co, KH, KW, Ci = w.shape
    x = torch.zeros((N, in_hw[0], in_hw[1], Ci), dtype=torch.float32, device=dy.device, requires_grad=True)
    y = _conv_f32(x, w, stride, pads)
    (dx,) = torch.autograd.grad(y, x, dy.float())
    if residual is not None:
        dx = dx + residual.float()
    return dx.to(BF).contiguous()


@torch.enable_grad()
def conv_wgrad(x, dy, dw, stride=1, pads=(0, 0, 0, 0), dbias=None):
    w = torch.zeros(dw.shape, dtype=torch.float32, device=x.device, requires_grad=True)
    y = _conv_f32(x, w, stride, pads)
    (g,) = torch.autograd.grad(y, w, dy.float())
    dw += g
    if dbias is not None:
        dbias += dy.float().reshape(-1, dy.shape[-1]).sum(0)


def _strided(t, sizes, strides):
    """View of t's storage starting at t's first element (t may itself be a column slice of a wider matrix)."""
    return torch.as_strided(t, sizes, strides, t.storage_offset())


def bgemm(a, a_desc, b, b_desc, c, c_desc, M, N, K, Z1, Z2, alpha=1.0, out_mode=0):
    ak, ald, as1, as2 = a_desc
    bk, bld, bs1, bs2 = b_desc
    A = _strided(a, (Z1, Z2, M, K), (as1, as2, ald, 1)) if ak else _strided(a, (Z1, Z2, K, M), (as1, as2, ald, 1)).transpose(-1, -2)
    B = _strided(b, (Z1, Z2, N, K), (bs1, bs2, bld, 1)) if bk else _strided(b, (Z1, Z2, K, N), (bs1, bs2, bld, 1)).transpose(-1, -2)
    r = alpha * (A.float() @ B.float().transpose(-1, -2))
    C = _strided(c, (Z1, Z2, M, N), (c_desc[1], c_desc[2], c_desc[0], 1))
    if out_mode == 2:
        C += r
    else:
        C.copy_(r.to(c.dtype))


# ---------------------------------------------------------------------------------------------- norms
def _gn_apply(x3, gamma, beta, G, eps, silu):
    S, P, C = x3.shape
    xf = x3.float().view(S, P, G, C // G)
    mean = xf.mean(dim=(1, 3), keepdim=True)
    var = xf.var(dim=(1, 3), unbiased=False, keepdim=True)
    rstd = (var + eps).rsqrt()
    y = ((xf - mean) * rstd).view(S, P, C) * gamma + beta
    if silu:
        y = F.silu(y)
    return y, mean.view(S, G), rstd.view(S, G)


def groupnorm_fwd(x, gamma, beta, G, eps, silu, stats=None, fps=1):
    S, P, C = x.shape
    if stats:   # the producer's sums ARE the statistics (as in the kernel): a wiring mistake upstream shows up in y
        st = torch.cat([t.view(S, fps, t.shape[1], 2).sum(dim=1) for t in stats], dim=1).double()      # [S, C, 2]
        cpg = C // G
        n = float(P * cpg)
        gsum = st.view(S, G, cpg, 2).sum(dim=2)
        mean = gsum[..., 0] / n
        var = (gsum[..., 1] / n - mean * mean).clamp_min(0)
        rstd = (1.0 / torch.sqrt(var + eps)).float()
        mean = mean.float()
        xh = (x.float().view(S, P, G, cpg) - mean[:, None, :, None]) * rstd[:, None, :, None]
        y = xh.view(S, P, C) * gamma.float() + beta.float()
        if silu:
            y = y * torch.sigmoid(y)
    else:
        y, mean, rstd = _gn_apply(x, gamma.float(), beta.float(), G, eps, silu)
    stat = torch.stack([mean, rstd], dim=-1).contiguous()
    cpg = C // G
    a = rstd.repeat_interleave(cpg, dim=1) * gamma.float()
    b = beta.float() - mean.repeat_interleave(cpg, dim=1) * a
    return y.to(BF).contiguous(), stat, torch.stack([a, b], dim=-1).contiguous()


def groupnorm_bwd(dy, x, gamma, stat, ab, G, silu, add=None, dgamma=None, dbeta=None):
    S, P, C = x.shape
    cpg = C // G
    mean = stat[..., 0].repeat_interleave(cpg, dim=1)[:, None, :]
    rstd = stat[..., 1].repeat_interleave(cpg, dim=1)[:, None, :]
    xh = (x.float() - mean) * rstd
    dz = dy.float()
    if silu:
        z = ab[..., 0][:, None, :] * x.float() + ab[..., 1][:, None, :]
        sg = torch.sigmoid(z)
        dz = dz * sg * (1 + z * (1 - sg))
    if dgamma is not None:
        dgamma += (dz * xh).sum(dim=(0, 1))
    if dbeta is not None:
        dbeta += dz.sum(dim=(0, 1))
    dxh = (dz * gamma.float()).view(S, P, G, cpg)
    xg = xh.view(S, P, G, cpg)
    m1 = dxh.mean(dim=(1, 3), keepdim=True)
    m2 = (dxh * xg).mean(dim=(1, 3), keepdim=True)
    dx = (rstd.view(S, 1, G, cpg) * (dxh - m1 - xg * m2)).view(S, P, C)
    if add is not None:
        dx = dx + add.float()
    return dx.to(BF).contiguous()


def layernorm_fwd(x, gamma, beta, eps):
    xf = x.float()
    mean = xf.mean(-1, keepdim=True)
    var = xf.var(-1, unbiased=False, keepdim=True)
    rstd = (var + eps).rsqrt()
    y = (xf - mean) * rstd * gamma.float() + beta.float()
    return y.to(BF).contiguous(), torch.cat([mean, rstd], dim=-1).contiguous()


def layernorm_bwd(dy, x, gamma, stat, add=None, dgamma=None, dbeta=None):
    xf = x.float()
    mean, rstd = stat[:, :1], stat[:, 1:]
    xh = (xf - mean) * rstd
    dyf = dy.float()
    dg = dyf * gamma.float()
    dx = rstd * (dg - dg.mean(-1, keepdim=True) - xh * (dg * xh).mean(-1, keepdim=True))
    if dgamma is not None:
        dgamma += (dyf * xh).sum(0)
    if dbeta is not None:
        dbeta += dyf.sum(0)
    if add is not None:
        dx = dx + add.float()
    return dx.to(BF).contiguous()


# ---------------------------------------------------------------------------------------------- elementwise / glue
def geglu_fwd(proj):
    h, g = proj.float().chunk(2, dim=-1)
    return (h * F.gelu(g)).to(BF).contiguous()


@torch.enable_grad()
def geglu_bwd(proj, dout):
    p = proj.float().requires_grad_(True)
    h, g = p.chunk(2, dim=-1)
    (dp,) = torch.autograd.grad(h * F.gelu(g), p, dout.float())
    return dp.to(BF).contiguous()


def silu_f32_to_bf16(x, apply_silu=True):
    return (F.silu(x) if apply_silu else x).to(BF)


@torch.enable_grad()
def silu_bwd_f32(x, dy):
    xx = x.clone().requires_grad_(True)
    (g,) = torch.autograd.grad(F.silu(xx), xx, dy)
    return g


def silu_bf16(x):
    return F.silu(x.float()).to(BF)


@torch.enable_grad()
def silu_bf16_bwd(x, dy):
    xx = x.float().requires_grad_(True)
    (g,) = torch.autograd.grad(F.silu(xx), xx, dy.float())
    return g.to(BF)


def add_bf16(a, b, c=None):
    r = a.float() + b.float()
    if c is not None:
        r = r + c.float()
    return r.to(BF)


def scale_bf16(a, alpha):
    return (a.float() * alpha).to(BF)


def _mix32(seed, idx):
    """splitmix64 finaliser, bit-identical to mix32() in elementwise.cu (int64 arithmetic wraps like uint64)."""
    def u64(v):
        v &= 0xFFFFFFFFFFFFFFFF
        return v - (1 << 64) if v >= (1 << 63) else v
    z = idx * u64(0x9E3779B97F4A7C15) + u64(seed)
    def lsr(v, s):
        return (v >> s) & ((1 << (64 - s)) - 1)
    z = (z ^ lsr(z, 30)) * u64(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * u64(0x94D049BB133111EB)
    return lsr(z ^ lsr(z, 31), 32)


def dropout_scale_add(x, base, p, scale, seed, epoch=None):
    if epoch is not None:   # same mixing as the kernel: seed ^= epoch * 0xD1342543DE82EF95 (mod 2^64)
        seed = (int(seed) ^ ((int(epoch.item()) * 0xD1342543DE82EF95) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF
    idx = torch.arange(x.numel(), dtype=torch.int64, device=x.device)
    keep = (_mix32(seed, idx) >= int(float(p) * 4294967296.0)).view(x.shape)
    k = torch.tensor(scale, dtype=torch.float32) / (torch.tensor(1.0, dtype=torch.float32) - torch.tensor(p, dtype=torch.float32))
    y = torch.where(keep, x.float() * k.to(x.device), torch.zeros((), device=x.device))
    if base is not None:
        y = y + base.float()
    return y.to(BF)


def counter_add(counter, value=1):
    counter.add_(value)


@torch.no_grad()
def sqnorm_chunks(g, chunks, out, g_bf16=None):
    src = g if g_bf16 is None else g_bf16
    for off, n in chunks.tolist():
        out[0] += src[off:off + n].double().pow(2).sum()


def scale_cast_f32_bf16(src, dst, alpha):
    dst.copy_((src * alpha).to(dst.dtype))


def cast_bf16_f32(src, dst):
    dst.copy_(src.float())


@torch.no_grad()
def adamw_prepare(hp_in, hp, state, sq, max_norm):
    """csrc/optim.cu adamw_prepare_kernel: step count, bias corrections, clip factor - on the 'device' tensors."""
    state[0] += 1
    step = int(state[0])
    norm = float(sq[0]) ** 0.5
    sq[1] = norm
    sq[0] = 0.0
    scale = min(1.0, max_norm / (norm + 1e-6)) if (max_norm or 0.0) > 0 else 1.0
    for s in range(hp_in.shape[0]):
        b1, b2 = float(hp_in[s, 1]), float(hp_in[s, 2])
        hp[s, :5] = hp_in[s]
        hp[s, 5] = 1.0 - b1 ** step
        hp[s, 6] = (1.0 - b2 ** step) ** 0.5
        hp[s, 7] = scale


@torch.no_grad()
def adamw_chunks(p, g, m, v, shadow, n_shadow, chunks, hp_row, zero_grad=True, g_bf16=None):
    """torch.optim.AdamW's single-tensor update on the chunk table (same arithmetic order as the kernel)."""
    lr, b1, b2, eps, wd, bc1, bc2s, gscale = (float(x) for x in hp_row.tolist())
    for off, n in chunks.tolist():
        sl = slice(off, off + n)
        gs = (g[sl] if g_bf16 is None else g_bf16[sl].float()) * gscale
        p[sl].mul_(1.0 - lr * wd)
        m[sl].mul_(b1).add_(gs, alpha=1.0 - b1)
        v[sl].mul_(b2).addcmul_(gs, gs, value=1.0 - b2)
        p[sl].addcdiv_(m[sl], v[sl].sqrt() / bc2s + eps, value=-(lr / bc1))
        if shadow is not None and off < n_shadow:
            shadow[sl].copy_(p[sl])
        if zero_grad:
            g[sl].zero_()


def add_f32(a, b):
    return a + b


def cast_f32_bf16(src, dst=None):
    if dst is None:
        return src.to(BF)
    dst.copy_(src)
    return dst


def upsample_nearest_fwd(x, out_hw_):
    return F.interpolate(x.float().permute(0, 3, 1, 2), size=tuple(out_hw_), mode="nearest").permute(0, 2, 3, 1).to(BF).contiguous()


@torch.enable_grad()
def upsample_nearest_bwd(dy, in_hw):
    N, Ho, Wo, C = dy.shape
    x = torch.zeros((N, C, in_hw[0], in_hw[1]), dtype=torch.float32, device=dy.device, requires_grad=True)
    y = F.interpolate(x, size=(Ho, Wo), mode="nearest")
    (g,) = torch.autograd.grad(y, x, dy.float().permute(0, 3, 1, 2))
    return g.permute(0, 2, 3, 1).to(BF).contiguous()


def concat_channels(a, b):
    return torch.cat([a, b], dim=-1).contiguous()


def split_channels(g, Ca):
    return g[..., :Ca].contiguous(), g[..., Ca:].contiguous()


def colsum(x, out, S, P, C):
    out += x.float().reshape(S, P, C).sum(1)


def colsum_f32(x, out):
    out += x.sum(0)


def frames_u8_to_nhwc8(frames, out_hw):
    x = frames.permute(0, 3, 1, 2).float()
    if tuple(x.shape[-2:]) != tuple(out_hw):
        x = F.interpolate(x, size=tuple(out_hw), mode="bilinear", align_corners=False)
    x = (x / 127.5 - 1.0).permute(0, 2, 3, 1)
    return torch.cat([x, torch.zeros(x.shape[:-1] + (5,), device=x.device)], dim=-1).to(BF).contiguous()


def gelu_bf16(x, quick=False):
    xf = x.float()
    return (xf * torch.sigmoid(1.702 * xf) if quick else F.gelu(xf)).to(BF)


def embed_tokens(ids, tok_emb, pos_emb):
    B, Lq = ids.shape
    return (tok_emb[ids] + pos_emb[:Lq][None]).reshape(B * Lq, -1).to(BF)


def softmax_fwd(s, n_valid, ld_out, causal_period=0):
    sc = s[..., :n_valid].float()
    if causal_period:
        rows = torch.arange(sc.numel() // sc.shape[-1], device=s.device).view(sc.shape[:-1]) % causal_period
        sc = sc.masked_fill(torch.arange(n_valid, device=s.device) > rows[..., None], float("-inf"))
    p = torch.softmax(sc, dim=-1)
    out = torch.zeros(s.shape[:-1] + (ld_out,), dtype=BF, device=s.device)
    out[..., :n_valid] = p.to(BF)
    return out


def softmax_bwd(p, dp, n_valid, scale):
    pf = p.float()[..., :n_valid]
    d = dp[..., :n_valid]
    ds = pf * (d - (pf * d).sum(-1, keepdim=True)) * scale
    out = torch.zeros_like(p)
    out[..., :n_valid] = ds.to(BF)
    return out


def _seq_view(t, addr, out_side):
    nseq, inner, outer_rows, inner_rows, seq_rows, ld_in, ld_out, heads, L, D = addr
    ld = ld_out if out_side else ld_in
    return torch.as_strided(t, (nseq // inner, inner, heads, L, D), (outer_rows * ld, inner_rows * ld, D, seq_rows * ld, 1),
                            t.storage_offset())


def attn_small_fwd(q, k, v, o, addr):
    D = addr[-1]
    Q, K, V = (_seq_view(t, addr, False).float() for t in (q, k, v))
    P = torch.softmax(Q @ K.transpose(-1, -2) * D ** -0.5, dim=-1)
    _seq_view(o, addr, True).copy_((P @ V).to(o.dtype))
    return o


@torch.enable_grad()
def attn_small_bwd(q, k, v, do, dq, dk, dv, addr):
    D = addr[-1]
    Q, K, V = (_seq_view(t, addr, False).float().detach().requires_grad_(True) for t in (q, k, v))
    O = torch.softmax(Q @ K.transpose(-1, -2) * D ** -0.5, dim=-1) @ V
    gq, gk, gv = torch.autograd.grad(O, (Q, K, V), _seq_view(do, addr, True).float())
    for g, dst in ((gq, dq), (gk, dk), (gv, dv)):
        _seq_view(dst, addr, False).copy_(g.to(dst.dtype))
    return dq, dk, dv


def _heads_view(t, heads):
    Nb, L, C = t.shape
    return t.float().reshape(Nb, L, heads, C // heads).transpose(1, 2)


def flash_attn_fwd(q, k, v, heads):
    """softmax(q k^T / sqrt(d)) v per (batch, head) and the log-sum-exp of the scaled scores (natural log)."""
    Nb, Lq, C = q.shape
    s = _heads_view(q, heads) @ _heads_view(k, heads).transpose(-1, -2) * (C // heads) ** -0.5
    o = (torch.softmax(s, -1) @ _heads_view(v, heads)).transpose(1, 2).reshape(Nb, Lq, C)
    return o.to(BF), torch.logsumexp(s, -1)


@torch.enable_grad()
def flash_attn_bwd(q, k, v, o, do, lse, heads, dq, dk, dv):
    Nb, Lq, C = q.shape
    qf, kf, vf = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
    s = _heads_view(qf, heads) @ _heads_view(kf, heads).transpose(-1, -2) * (C // heads) ** -0.5
    out = (torch.softmax(s, -1) @ _heads_view(vf, heads)).transpose(1, 2).reshape(Nb, Lq, C)
    gq, gk, gv = torch.autograd.grad(out, (qf, kf, vf), do.float())
    for g, dst in ((gq, dq), (gk, dk), (gv, dv)):
        dst.copy_(g.to(dst.dtype))


def timestep_embedding(t, dim):
    half = dim // 2
    f = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    a = t[:, None].float() * f[None]
    return torch.cat([torch.cos(a), torch.sin(a)], dim=-1).to(BF)


def latents_to_nhwc8(x0, noise=None, alphas_cumprod=None, timesteps=None):
    B, C, Fr, H, W = x0.shape
    x = x0
    if noise is not None:
        a = alphas_cumprod[timesteps].view(B, 1, 1, 1, 1)
        x = a.sqrt() * x0 + (1 - a).sqrt() * noise
    out = torch.zeros((B, Fr, H, W, 8), dtype=torch.float32, device=x0.device)
    out[..., :C] = x.permute(0, 2, 3, 4, 1)
    return out.view(B * Fr, H, W, 8).to(BF)


def nhwc8_to_latents(x, B, C, Fr):
    _, H, W, _ = x.shape
    return x.float().view(B, Fr, H, W, 8)[..., :C].permute(0, 4, 1, 2, 3).contiguous()


def vae_sample(moments, eps, B, Fr, scale):
    _, h, w, _ = moments.shape
    m = moments.float().view(B, Fr, h, w, 8)
    mean, logvar = m[..., :4].permute(0, 4, 1, 2, 3), m[..., 4:].permute(0, 4, 1, 2, 3).clamp(-30.0, 20.0)
    return ((mean + torch.exp(0.5 * logvar) * eps) * scale).contiguous()


def mse_loss_fwd(pred, target):
    B, C, Fr, H, W = target.shape
    p = nhwc8_to_latents(pred, B, C, Fr)
    return ((p - target) ** 2).mean()


def mse_loss_bwd(pred, target, gout):
    B, C, Fr, H, W = target.shape
    p = nhwc8_to_latents(pred, B, C, Fr)
    g = 2.0 * (p - target) / p.numel() * gout
    return latents_to_nhwc8(g)


ALL = [n for n, f in list(globals().items()) if callable(f) and not n.startswith("_") and n not in ("F",)]
