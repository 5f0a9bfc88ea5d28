"""Stand-in for diffusers.models.transformer_temporal as a plain nn.Module; independent of oracle/leaves.py
(see embeddings.py)."""
from dataclasses import dataclass

import torch
import torch.nn as nn

from ..utils import BaseOutput
from .attention import BasicTransformerBlock


@dataclass
class TransformerTemporalModelOutput(BaseOutput):
    sample: torch.FloatTensor


class TransformerTemporalModel(nn.Module):
    def __init__(self, num_attention_heads=16, attention_head_dim=88, in_channels=None, out_channels=None, num_layers=1,
                 dropout=0.0, norm_num_groups=32, cross_attention_dim=None, attention_bias=False, sample_size=None,
                 activation_fn="geglu", norm_elementwise_affine=True, double_self_attention=True):
        super().__init__()
        assert num_layers == 1 and double_self_attention
        inner = num_attention_heads * attention_head_dim
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, cross_attention_dim, double_self_attention=True)])
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, class_labels=None, num_frames=1,
                cross_attention_kwargs=None, return_dict=True):
        batch_frames, channel, height, width = hidden_states.shape
        batch_size = batch_frames // num_frames
        residual = hidden_states
        hidden_states = hidden_states[None, :].reshape(batch_size, num_frames, channel, height, width).permute(0, 2, 1, 3, 4)
        hidden_states = self.norm(hidden_states)   # GroupNorm over (frames, height, width) of each clip
        hidden_states = hidden_states.permute(0, 3, 4, 2, 1).reshape(batch_size * height * width, num_frames, channel)
        hidden_states = self.proj_in(hidden_states)
        for block in self.transformer_blocks:
            hidden_states = block(hidden_states, encoder_hidden_states=encoder_hidden_states)
        hidden_states = self.proj_out(hidden_states)
        hidden_states = (hidden_states[None, None, :].reshape(batch_size, height, width, num_frames, channel)
                         .permute(0, 3, 4, 1, 2).contiguous())
        hidden_states = hidden_states.reshape(batch_frames, channel, height, width)
        out = hidden_states + residual
        return TransformerTemporalModelOutput(sample=out) if return_dict else (out,)
