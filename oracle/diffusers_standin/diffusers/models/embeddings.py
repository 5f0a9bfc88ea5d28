"""Stand-in for diffusers.models.embeddings (test infrastructure; the real package cannot be installed offline).

Written as ordinary nn.Modules on torch's own operators - it does NOT call oracle/leaves.py.  The reference's
models/*.py run unmodified on top of these classes; tests/test_oracle_vs_reference.py then compares that stack with the
functional restatement in oracle/leaves.py + oracle/unet3d_ref.py: two independent write-ups of the diffusers leaf
semantics that must agree to fp32 round-off."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class Timesteps(nn.Module):
    """get_timestep_embedding: sin | cos halves of t * 10000^(-i / (half - shift)), flipped to cos | sin on request."""

    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
        super().__init__()
        self.num_channels, self.flip, self.shift = num_channels, flip_sin_to_cos, downscale_freq_shift

    def forward(self, timesteps):
        half = self.num_channels // 2
        exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / (half - self.shift)
        emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        if self.flip:
            emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
        if self.num_channels % 2 == 1:
            emb = F.pad(emb, (0, 1, 0, 0))
        return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim, act_fn="silu", out_dim=None, post_act_fn=None, cond_proj_dim=None):
        super().__init__()
        assert act_fn == "silu" and cond_proj_dim is None and post_act_fn is None
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, out_dim or time_embed_dim)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))
