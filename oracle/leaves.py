"""ORACLE (test infrastructure, not product code): CPU restatement of the leaf modules the reference delegates to.

The reference (ExponentialML/Text-To-Video-Finetuning) owns only the *wiring* of its UNet
(models/unet_3d_condition.py, models/unet_3d_blocks.py); every leaf it instantiates comes from the un-vendored,
un-pinned `diffusers` dependency (requirements.txt:5, `git+https://github.com/huggingface/diffusers.git`, no tag;
the import paths at unet_3d_blocks.py:18-20 and train.py:35-36 bracket it to ~0.15-0.25).  diffusers is absent from
/root/reference and from this image, so its published algorithm is restated here in plain fp32/fp64 torch.

PARITY UNPINNED: the reference ships no tests, golden vectors or known-answer values for this path (SURVEY.md
section 4 / 8c).  What *is* pinned: (a) tests/test_oracle_vs_reference.py imports the reference's own
models/*.py unmodified on top of oracle/diffusers_standin (which calls the functions below) and checks this
restatement's wiring against it; (b) structural constants (1,411,233,860 UNet parameters, 1,480 tensors,
34,163,664 VAE-encoder parameters).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this package.

All functions are functional: `p` is a mapping name -> tensor (a state dict or dict(module.named_parameters()))
and `pre` the key prefix of the module being evaluated.  Layout is the reference's (N, C, H, W).
"""
import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ embeddings
def timestep_sinusoid(timesteps, dim=320, flip_sin_to_cos=True, downscale_freq_shift=0.0, max_period=10000):
    """diffusers Timesteps / get_timestep_embedding as configured at unet_3d_condition.py:138
    (Timesteps(block_out_channels[0], True, 0))."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


def timestep_embedding(p, pre, t_emb):
    """TimestepEmbedding(320 -> 1280, act 'silu') - unet_3d_condition.py:141-145."""
    h = F.linear(t_emb, p[pre + "linear_1.weight"], p[pre + "linear_1.bias"])
    return F.linear(F.silu(h), p[pre + "linear_2.weight"], p[pre + "linear_2.bias"])


# ------------------------------------------------------------------------------------------------ resnet family
def resnet_block2d(p, pre, x, temb, groups=32, eps=1e-5, output_scale_factor=1.0):
    """ResnetBlock2D (pre-norm, 'default' time embedding norm, swish, dropout 0) as built at
    unet_3d_blocks.py:295-306,457-469,597-608,693-704,827-838.  GroupNorm is per frame."""
    h = F.group_norm(x, groups, p[pre + "norm1.weight"], p[pre + "norm1.bias"], eps)
    h = F.conv2d(F.silu(h), p[pre + "conv1.weight"], p[pre + "conv1.bias"], padding=1)
    if temb is not None and (pre + "time_emb_proj.weight") in p:
        t = F.linear(F.silu(temb), p[pre + "time_emb_proj.weight"], p[pre + "time_emb_proj.bias"])
        h = h + t[:, :, None, None]
    h = F.group_norm(h, groups, p[pre + "norm2.weight"], p[pre + "norm2.bias"], eps)
    h = F.conv2d(F.silu(h), p[pre + "conv2.weight"], p[pre + "conv2.bias"], padding=1)
    if (pre + "conv_shortcut.weight") in p:
        x = F.conv2d(x, p[pre + "conv_shortcut.weight"], p[pre + "conv_shortcut.bias"])
    return (x + h) / output_scale_factor


def temporal_conv_layer(p, pre, x, num_frames, groups=32, eps=1e-5):
    """TemporalConvLayer(C, C, dropout=0.1) in eval/p=0 mode (unet_3d_blocks.py:308-314): four
    [GroupNorm(32) over (C/32, F, H, W) -> SiLU -> (Dropout) -> Conv3d (3,1,1) pad (1,0,0)] + identity."""
    bf, c, hh, ww = x.shape
    h = x[None, :].reshape((-1, num_frames) + x.shape[1:]).permute(0, 2, 1, 3, 4)
    identity = h
    for name, conv_idx in (("conv1", 2), ("conv2", 3), ("conv3", 3), ("conv4", 3)):
        h = F.group_norm(h, groups, p[f"{pre}{name}.0.weight"], p[f"{pre}{name}.0.bias"], eps)
        h = F.conv3d(F.silu(h), p[f"{pre}{name}.{conv_idx}.weight"], p[f"{pre}{name}.{conv_idx}.bias"], padding=(1, 0, 0))
    h = identity + h
    return h.permute(0, 2, 1, 3, 4).reshape(bf, c, hh, ww)


def downsample2d(p, pre, x, padding=1):
    """Downsample2D(use_conv=True): 3x3 stride-2 conv; the VAE variant (padding=0) pads (0,1,0,1) first."""
    if padding == 0:
        x = F.pad(x, (0, 1, 0, 1))
    return F.conv2d(x, p[pre + "conv.weight"], p[pre + "conv.bias"], stride=2, padding=padding)


def upsample2d(p, pre, x, output_size=None):
    """Upsample2D(use_conv=True): nearest x2 (or explicit size) then 3x3 conv."""
    if output_size is None:
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
    else:
        x = F.interpolate(x, size=output_size, mode="nearest")
    return F.conv2d(x, p[pre + "conv.weight"], p[pre + "conv.bias"], padding=1)


# ------------------------------------------------------------------------------------------------ attention family
def attention(p, pre, x, context, heads):
    """diffusers Attention (bias-free q/k/v, biased to_out[0], no mask, scale d^-0.5)."""
    ctx = x if context is None else context
    q = F.linear(x, p[pre + "to_q.weight"], p.get(pre + "to_q.bias"))
    k = F.linear(ctx, p[pre + "to_k.weight"], p.get(pre + "to_k.bias"))
    v = F.linear(ctx, p[pre + "to_v.weight"], p.get(pre + "to_v.bias"))
    b, lq, inner = q.shape
    d = inner // heads
    q = q.view(b, lq, heads, d).transpose(1, 2)
    k = k.view(b, -1, heads, d).transpose(1, 2)
    v = v.view(b, -1, heads, d).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
    o = torch.matmul(torch.softmax(s, dim=-1), v)
    o = o.transpose(1, 2).reshape(b, lq, inner)
    return F.linear(o, p[pre + "to_out.0.weight"], p[pre + "to_out.0.bias"])


def feed_forward_geglu(p, pre, x):
    """FeedForward(dim, mult 4, 'geglu'): net.0 = GEGLU(proj: dim -> 8 dim; h * gelu(gate)), net.2 = Linear(4 dim -> dim)."""
    h, gate = F.linear(x, p[pre + "net.0.proj.weight"], p[pre + "net.0.proj.bias"]).chunk(2, dim=-1)
    return F.linear(h * F.gelu(gate), p[pre + "net.2.weight"], p[pre + "net.2.bias"])


def basic_transformer_block(p, pre, x, context, heads, double_self_attention=False):
    """BasicTransformerBlock: x += attn1(LN1 x); x += attn2(LN2 x, ctx); x += ff(LN3 x); LayerNorm eps 1e-5."""
    c = x.shape[-1]
    n = F.layer_norm(x, (c,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], 1e-5)
    x = attention(p, pre + "attn1.", n, None, heads) + x
    n = F.layer_norm(x, (c,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], 1e-5)
    x = attention(p, pre + "attn2.", n, None if double_self_attention else context, heads) + x
    n = F.layer_norm(x, (c,), p[pre + "norm3.weight"], p[pre + "norm3.bias"], 1e-5)
    return feed_forward_geglu(p, pre + "ff.", n) + x


def transformer2d(p, pre, x, context, heads, groups=32):
    """Transformer2DModel(use_linear_projection=True, num_layers=1), unet_3d_blocks.py:319-330."""
    n_, c, hh, ww = x.shape
    h = F.group_norm(x, groups, p[pre + "norm.weight"], p[pre + "norm.bias"], 1e-6)
    h = h.permute(0, 2, 3, 1).reshape(n_, hh * ww, c)
    h = F.linear(h, p[pre + "proj_in.weight"], p[pre + "proj_in.bias"])
    h = basic_transformer_block(p, pre + "transformer_blocks.0.", h, context, heads)
    h = F.linear(h, p[pre + "proj_out.weight"], p[pre + "proj_out.bias"])
    h = h.reshape(n_, hh, ww, c).permute(0, 3, 1, 2)
    return h + x


def transformer_temporal(p, pre, x, num_frames, heads, groups=32):
    """TransformerTemporalModel(double_self_attention=True, num_layers=1): attention along the frame axis."""
    bf, c, hh, ww = x.shape
    b = bf // num_frames
    h = x[None, :].reshape(b, num_frames, c, hh, ww).permute(0, 2, 1, 3, 4)
    h = F.group_norm(h, groups, p[pre + "norm.weight"], p[pre + "norm.bias"], 1e-6)
    h = h.permute(0, 3, 4, 2, 1).reshape(b * hh * ww, num_frames, c)
    h = F.linear(h, p[pre + "proj_in.weight"], p[pre + "proj_in.bias"])
    h = basic_transformer_block(p, pre + "transformer_blocks.0.", h, None, heads, double_self_attention=True)
    h = F.linear(h, p[pre + "proj_out.weight"], p[pre + "proj_out.bias"])
    h = h[None, None, :].reshape(b, hh, ww, num_frames, c).permute(0, 3, 4, 1, 2).contiguous()
    return h.reshape(bf, c, hh, ww) + x


# ------------------------------------------------------------------------------------------------ VAE encoder
def vae_attention(p, pre, x, groups=32):
    """AutoencoderKL mid-block Attention: 1 head, d = C, GroupNorm(eps 1e-6), biased q/k/v/out, residual."""
    n_, c, hh, ww = x.shape
    h = F.group_norm(x, groups, p[pre + "group_norm.weight"], p[pre + "group_norm.bias"], 1e-6)
    h = h.view(n_, c, hh * ww).transpose(1, 2)
    h = attention(p, pre, h, None, heads=1)
    return h.transpose(1, 2).reshape(n_, c, hh, ww) + x


def vae_encode_moments(p, x, block_out_channels=(128, 256, 512, 512), layers_per_block=2, groups=32, pre=""):
    """AutoencoderKL.encode up to the Gaussian moments (SD-VAE encoder + quant_conv); train.py:339-347 calls
    vae.encode(t).latent_dist.sample() on this."""
    e = pre + "encoder."
    h = F.conv2d(x, p[e + "conv_in.weight"], p[e + "conv_in.bias"], padding=1)
    for i in range(len(block_out_channels)):
        for j in range(layers_per_block):
            h = resnet_block2d(p, f"{e}down_blocks.{i}.resnets.{j}.", h, None, groups, 1e-6)
        if i != len(block_out_channels) - 1:
            h = downsample2d(p, f"{e}down_blocks.{i}.downsamplers.0.", h, padding=0)
    h = resnet_block2d(p, e + "mid_block.resnets.0.", h, None, groups, 1e-6)
    h = vae_attention(p, e + "mid_block.attentions.0.", h, groups)
    h = resnet_block2d(p, e + "mid_block.resnets.1.", h, None, groups, 1e-6)
    h = F.group_norm(h, groups, p[e + "conv_norm_out.weight"], p[e + "conv_norm_out.bias"], 1e-6)
    h = F.conv2d(F.silu(h), p[e + "conv_out.weight"], p[e + "conv_out.bias"], padding=1)
    return F.conv2d(h, p[pre + "quant_conv.weight"], p[pre + "quant_conv.bias"])


def vae_decode(p, z, block_out_channels=(128, 256, 512, 512), layers_per_block=2, groups=32, pre=""):
    """AutoencoderKL.decode: post_quant_conv + SD-VAE decoder (mid block, 4 up blocks of layers_per_block + 1 resnets, nearest
    x2 upsample + conv between them).  z: unscaled latents (N, 4, h, w).  Used by the validation preview (train.py:908-958)."""
    d = pre + "decoder."
    rev = tuple(reversed(block_out_channels))
    h = F.conv2d(z, p[pre + "post_quant_conv.weight"], p[pre + "post_quant_conv.bias"])
    h = F.conv2d(h, p[d + "conv_in.weight"], p[d + "conv_in.bias"], padding=1)
    h = resnet_block2d(p, d + "mid_block.resnets.0.", h, None, groups, 1e-6)
    h = vae_attention(p, d + "mid_block.attentions.0.", h, groups)
    h = resnet_block2d(p, d + "mid_block.resnets.1.", h, None, groups, 1e-6)
    for i in range(len(rev)):
        for j in range(layers_per_block + 1):
            h = resnet_block2d(p, f"{d}up_blocks.{i}.resnets.{j}.", h, None, groups, 1e-6)
        if i != len(rev) - 1:
            h = upsample2d(p, f"{d}up_blocks.{i}.upsamplers.0.", h, None)
    h = F.group_norm(h, groups, p[d + "conv_norm_out.weight"], p[d + "conv_norm_out.bias"], 1e-6)
    return F.conv2d(F.silu(h), p[d + "conv_out.weight"], p[d + "conv_out.bias"], padding=1)


def diagonal_gaussian_sample(moments, eps_noise):
    """DiagonalGaussianDistribution.sample(): mean + exp(0.5 * clamp(logvar, -30, 20)) * eps."""
    mean, logvar = moments.chunk(2, dim=1)
    return mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * eps_noise


# ------------------------------------------------------------------------------------------------ scheduler
def ddpm_alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
    """DDPMScheduler 'scaled_linear' betas of the ms-1.7b / zeroscope scheduler config."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def add_noise(x0, noise, timesteps, alphas_cumprod):
    """DDPMScheduler.add_noise (train.py:760): sqrt(abar_t) x0 + sqrt(1 - abar_t) eps, abar broadcast per clip."""
    a = alphas_cumprod.to(x0.device)[timesteps].to(x0.dtype)
    sa, sb = a.sqrt(), (1 - a).sqrt()
    while sa.dim() < x0.dim():
        sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
    return sa * x0 + sb * noise
