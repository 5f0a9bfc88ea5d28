"""Stand-in for diffusers.models.attention / attention_processor (AttnProcessor2_0 semantics) as plain nn.Modules on
torch's own operators; independent of oracle/leaves.py (see embeddings.py)."""
import torch.nn as nn
import torch.nn.functional as F


class Attention(nn.Module):
    """q/k/v projections without bias, softmax(q k^T / sqrt(d)) v per head (F.scaled_dot_product_attention, what
    AttnProcessor2_0 calls), output projection with bias, dropout."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, bias=False):
        super().__init__()
        inner = heads * dim_head
        ctx = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(ctx, inner, bias=bias)
        self.to_v = nn.Linear(ctx, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def set_processor(self, processor):
        pass

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        B, L, _ = hidden_states.shape
        q, k, v = self.to_q(hidden_states), self.to_k(ctx), self.to_v(ctx)
        split = lambda t: t.view(B, -1, self.heads, t.shape[-1] // self.heads).transpose(1, 2)  # noqa: E731
        o = F.scaled_dot_product_attention(split(q), split(k), split(v), attn_mask=None, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(B, L, -1)
        return self.to_out[1](self.to_out[0](o))


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, hidden_states):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)   # exact (erf) GELU


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, hidden_states):
        for module in self.net:
            hidden_states = module(hidden_states)
        return hidden_states


class BasicTransformerBlock(nn.Module):
    """x += attn1(norm1 x); x += attn2(norm2 x, ctx); x += ff(norm3 x)  (pre-LayerNorm, eps 1e-5)."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, cross_attention_dim=None, double_self_attention=False):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, num_attention_heads, attention_head_dim)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, None if double_self_attention else cross_attention_dim, num_attention_heads, attention_head_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)
        self.double_self_attention = double_self_attention

    def forward(self, hidden_states, encoder_hidden_states=None):
        hidden_states = self.attn1(self.norm1(hidden_states)) + hidden_states
        ctx = None if self.double_self_attention else encoder_hidden_states
        hidden_states = self.attn2(self.norm2(hidden_states), encoder_hidden_states=ctx) + hidden_states
        return self.ff(self.norm3(hidden_states)) + hidden_states
