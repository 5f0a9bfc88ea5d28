from dataclasses import dataclass

import torch
import torch.nn as nn

from oracle import leaves as L
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
        self.heads, self.groups = num_attention_heads, norm_num_groups
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, cross_attention_dim, double_self_attention=True)])
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, class_labels=None, num_frames=1,
                cross_attention_kwargs=None, return_dict=True):
        out = L.transformer_temporal(dict(self.named_parameters()), "", hidden_states, num_frames, self.heads, self.groups)
        return TransformerTemporalModelOutput(sample=out) if return_dict else (out,)
