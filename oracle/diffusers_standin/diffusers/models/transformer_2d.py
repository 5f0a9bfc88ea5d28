from dataclasses import dataclass

import torch
import torch.nn as nn

from oracle import leaves as L
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
        self.heads, self.groups = num_attention_heads, norm_num_groups
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, cross_attention_dim)])
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, class_labels=None,
                cross_attention_kwargs=None, attention_mask=None, return_dict=True):
        out = L.transformer2d(dict(self.named_parameters()), "", hidden_states, encoder_hidden_states, self.heads, self.groups)
        return Transformer2DModelOutput(sample=out) if return_dict else (out,)
