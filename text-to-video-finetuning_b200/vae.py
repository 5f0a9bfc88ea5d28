"""AutoencoderKL on the B200-native kernels - SURVEY section 8 row a1 (`tensor_to_vae_latent`, train.py:339-347) and, for the
validation preview (8(f) row 4, train.py:908-958), the decoder.

Same parameter names as the diffusers SD-VAE (encoder.* / quant_conv.* / decoder.* / post_quant_conv.*), same call surface
train.py uses: `vae.encode(x).latent_dist.sample()`, `vae.decode(z).sample`, `enable_slicing()`, `.to()`, `.dtype`.  The
decoder is built on request (`build_decoder=True`, or automatically when a checkpoint carries decoder weights): training
needs the encoder only (34,163,664 parameters), the preview sampler the decoder too.
Differences by design: all frames are encoded as ONE batch (the reference's `enable_slicing` encodes frame by frame,
H12) on channels-last bf16 activations; the mid-block attention (1 head, d = 512) runs as batched tcgen05 GEMMs; the
Gaussian sample, the (b f) c h w -> b c f h w rearrange and the 0.18215 scale are one kernel (`t2v_vae_sample`).
"""
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import ops, prims
from .layers import Attention, Downsample2D, ResnetBlock2D, Upsample2D, _channels_last_, run_conv, run_group_norm, run_linear
from .modeling_utils import ConfigMixin, ModelMixin, register_to_config


class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_downsample, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels=in_channels if i == 0 else out_channels, out_channels=out_channels,
                                                    temb_channels=None, groups=groups, eps=eps) for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels, padding=0, name="op")]) \
            if add_downsample else None

    def forward(self, h):
        for r in self.resnets:
            h = r(h)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
        return h


class VaeAttention(Attention):
    """Single-head self-attention over the h*w tokens of a frame, with its own GroupNorm and residual."""

    def __init__(self, channels, groups, eps):
        super().__init__(channels, None, heads=1, dim_head=channels, bias=True)
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps)

    def forward(self, x):
        N, H, W, C = x.shape
        res, h = ops.fork(x)
        n = run_group_norm(self.group_norm, h, False, N).view(N * H * W, C)
        n1, n2, n3 = ops.fork(n, 3)
        q, k, v = run_linear(self.to_q, n1), run_linear(self.to_k, n2), run_linear(self.to_v, n3)
        a = ops.attention(q.view(N, H * W, C), k.view(N, H * W, C), v.view(N, H * W, C), 1).view(N * H * W, C)
        return self.project_out(a, res.view(N * H * W, C)).view(N, H, W, C)


class _MidBlock(nn.Module):
    def __init__(self, channels, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels=channels, out_channels=channels, temb_channels=None, groups=groups, eps=eps)
                                      for _ in range(2)])
        self.attentions = nn.ModuleList([VaeAttention(channels, groups, eps)])

    def forward(self, h):
        return self.resnets[1](self.attentions[0](self.resnets[0](h)))


class Encoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups, eps=1e-6):
        super().__init__()
        self.in_channels = in_channels
        self.conv_in = _channels_last_(nn.Conv2d(in_channels, block_out_channels[0], 3, padding=1))
        blocks, ch = [], block_out_channels[0]
        for i, oc in enumerate(block_out_channels):
            blocks.append(DownEncoderBlock2D(ch, oc, layers_per_block, i != len(block_out_channels) - 1, norm_num_groups, eps))
            ch = oc
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = _MidBlock(ch, norm_num_groups, eps)
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, ch, eps=eps)
        self.conv_act = nn.SiLU()
        self.conv_out = _channels_last_(nn.Conv2d(ch, 2 * out_channels, 3, padding=1))

    def forward(self, x):
        h = run_conv(self.conv_in, x)
        for b in self.down_blocks:
            h = b(h)
        h = self.mid_block(h)
        h = run_group_norm(self.conv_norm_out, h, True, h.shape[0])
        return run_conv(self.conv_out, h)


class UpDecoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_upsample, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels=in_channels if i == 0 else out_channels, out_channels=out_channels,
                                                    temb_channels=None, groups=groups, eps=eps) for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)]) if add_upsample else None

    def forward(self, h):
        for r in self.resnets:
            h = r(h)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


class Decoder(nn.Module):
    """SD-VAE decoder: conv_in - mid (resnet, 1-head attention, resnet) - 4 up blocks of layers_per_block + 1 resnets with
    nearest x2 upsampling between them - GroupNorm / SiLU / conv_out."""

    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups, eps=1e-6):
        super().__init__()
        self.out_channels = out_channels
        rev = tuple(reversed(block_out_channels))
        self.conv_in = _channels_last_(nn.Conv2d(in_channels, rev[0], 3, padding=1))
        self.mid_block = _MidBlock(rev[0], norm_num_groups, eps)
        blocks, ch = [], rev[0]
        for i, oc in enumerate(rev):
            blocks.append(UpDecoderBlock2D(ch, oc, layers_per_block + 1, i != len(rev) - 1, norm_num_groups, eps))
            ch = oc
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, ch, eps=eps)
        self.conv_act = nn.SiLU()
        self.conv_out = _channels_last_(nn.Conv2d(ch, out_channels, 3, padding=1))

    def forward(self, z):
        h = run_conv(self.conv_in, z, cin_pad=8 - self.conv_in.in_channels)
        h = self.mid_block(h)
        for b in self.up_blocks:
            h = b(h)
        h = run_group_norm(self.conv_norm_out, h, True, h.shape[0])
        return run_conv(self.conv_out, h, cout_pad=8 - self.out_channels)


class LatentDist:
    """Stand-in for diffusers' DiagonalGaussianDistribution holding the channels-last moments."""

    def __init__(self, moments_nhwc8, n):
        self.moments, self.n = moments_nhwc8, n

    def sample(self, generator=None):
        _, h, w, _ = self.moments.shape
        eps = torch.randn((self.n, 4, 1, h, w), device=self.moments.device, dtype=torch.float32, generator=generator)
        return prims.vae_sample(self.moments, eps, self.n, 1, 1.0).view(self.n, 4, h, w)

    def mode(self):
        _, h, w, _ = self.moments.shape
        return prims.nhwc8_to_latents(self.moments, self.n, 4, 1).view(self.n, 4, h, w)


class AutoencoderKL(ModelMixin, ConfigMixin):
    @register_to_config
    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2, latent_channels=4,
                 norm_num_groups=32, scaling_factor=0.18215, build_decoder=False, **unused):
        super().__init__()
        if latent_channels != 4 or in_channels > 8:
            raise NotImplementedError("only the SD-VAE (3 -> 4 latent channels) is implemented")
        self.encoder = Encoder(in_channels, latent_channels, tuple(block_out_channels), layers_per_block, norm_num_groups)
        self.quant_conv = _channels_last_(nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1))
        self.decoder = self.post_quant_conv = None
        self._dec_args = (latent_channels, out_channels, tuple(block_out_channels), layers_per_block, norm_num_groups)
        if build_decoder:
            self.build_decoder()
        self.use_slicing = False

    def build_decoder(self):
        if self.decoder is None:
            lc, oc, boc, lpb, g = self._dec_args
            dev = self.quant_conv.weight.device
            self.decoder = Decoder(lc, oc, boc, lpb, g).to(dev)
            self.post_quant_conv = _channels_last_(nn.Conv2d(lc, lc, 1)).to(dev)
        return self

    def enable_slicing(self):
        """Accepted for train.py:678 compatibility.  Frames are always encoded as one batch here (H12)."""
        self.use_slicing = True

    def disable_slicing(self):
        self.use_slicing = False

    def load_state_dict(self, state_dict, strict=True):
        """Accepts full diffusers VAE checkpoints: decoder / post_quant_conv tensors are ignored, and the pre-0.15
        attention key names (query/key/value/proj_attn) are mapped to to_q/to_k/to_v/to_out.0."""
        ren = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}
        has_dec = any(k.startswith("decoder.") for k in state_dict)
        if has_dec:
            self.build_decoder()   # a full checkpoint brings the decoder along (validation preview)
        sd = {}
        for k, v in state_dict.items():
            if (k.startswith("decoder.") or k.startswith("post_quant_conv.")) and self.decoder is None:
                continue
            parts = k.split(".")
            if "attentions" in parts and parts[-2] in ren:
                parts[-2:-1] = ren[parts[-2]].split(".")
                k = ".".join(parts)
            sd[k] = v
        return super().load_state_dict(sd, strict=strict)

    @torch.no_grad()
    def encode_moments(self, x):
        """x (N, 3, H, W) in [-1, 1] -> moments [N, H/8, W/8, 8] bf16 channels-last (mean | logvar)."""
        N, C, H, W = x.shape
        return self.encode_moments_nhwc8(prims.latents_to_nhwc8(x.float().contiguous().view(N, C, 1, H, W)))

    @torch.no_grad()
    def encode_moments_nhwc8(self, xin):
        """xin bf16 [N, H, W, 8] channels-last in [-1, 1] (what prims.frames_u8_to_nhwc8 produces) -> moments [N, H/8, W/8, 8]."""
        return run_conv(self.quant_conv, self.encoder(xin), pads=(0, 0, 0, 0))

    def encode(self, x, return_dict=True):
        return SimpleNamespace(latent_dist=LatentDist(self.encode_moments(x), x.shape[0]))

    @torch.no_grad()
    def decode(self, z, return_dict=True):
        """z (N, 4, h, w) UNSCALED latents (the caller divides by 0.18215, train.py / pipeline decode_latents) -> `.sample`
        (N, 3, 8h, 8w) in about [-1, 1], fp32."""
        if self.decoder is None:
            raise RuntimeError("this AutoencoderKL was built without its decoder (build_decoder=True or load a full checkpoint)")
        N, C, h, w = z.shape
        zin = prims.latents_to_nhwc8(z.float().contiguous().view(N, C, 1, h, w))
        zin = run_conv(self.post_quant_conv, zin, pads=(0, 0, 0, 0), cin_pad=8 - C, cout_pad=8 - C)
        out = self.decoder(zin)
        img = prims.nhwc8_to_latents(out, N, 3, 1).view(N, 3, out.shape[1], out.shape[2])
        return SimpleNamespace(sample=img) if return_dict else (img,)


@torch.no_grad()
def tensor_to_vae_latent(t, vae, generator=None):
    """train.py:339-347: (B, F, 3, H, W) pixels -> (B, 4, F, H/8, W/8) latents * 0.18215, one batched encode + one
    fused sample/rearrange/scale kernel."""
    B, F = t.shape[:2]
    mom = vae.encode_moments(t.reshape((B * F,) + tuple(t.shape[2:])))
    _, h, w, _ = mom.shape
    eps = torch.randn((B, 4, F, h, w), device=mom.device, dtype=torch.float32, generator=generator)
    return prims.vae_sample(mom, eps, B, F, 0.18215)
