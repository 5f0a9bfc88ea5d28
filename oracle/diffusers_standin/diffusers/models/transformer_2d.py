"""Stand-in for diffusers.models.transformer_2d (linear-projection variant) as a plain nn.Module; independent of
oracle/leaves.py (see embeddings.py)."""
from dataclasses import dataclass

import torch
import torch.nn as nn

from ..utils import BaseOutput
from .attention import BasicTransformerBlock


@dataclass
class Transformer2DModelOutput(BaseOutput):
    sample: torch.FloatTensor


class Transformer2DModel(nn.Module):
    def __init__(self, num_attention_heads=16, attention_head_dim=88, in_channels=None, out_channels=None, num_layers=1,
                 dropout=0.0, norm_num_groups=32, cross_attention_dim=None, attention_bias=False, sample_size=None,
                 num_vector_embeds=None, patch_size=None, activation_fn="geglu", num_embeds_ada_norm=None,
                 use_linear_projection=False, only_cross_attention=False, upcast_attention=False, **unused):
        super().__init__()
        assert use_linear_projection and num_layers == 1 and not only_cross_attention
        inner = num_attention_heads * attention_head_dim
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, cross_attention_dim)])
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, class_labels=None,
                cross_attention_kwargs=None, attention_mask=None, return_dict=True):
        batch, channels, height, width = hidden_states.shape
        residual = hidden_states
        hidden_states = self.norm(hidden_states)
        hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, channels)
        hidden_states = self.proj_in(hidden_states)
        for block in self.transformer_blocks:
            hidden_states = block(hidden_states, encoder_hidden_states=encoder_hidden_states)
        hidden_states = self.proj_out(hidden_states)
        hidden_states = hidden_states.reshape(batch, height, width, channels).permute(0, 3, 1, 2).contiguous()
        out = hidden_states + residual
        return Transformer2DModelOutput(sample=out) if return_dict else (out,)
