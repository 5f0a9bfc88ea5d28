"""Frozen CLIP text encoder on the B200-native kernels - SURVEY 8(f) row 2 (`text_encoder(token_ids)[0]`, train.py:784-790).

Drop-in for the `transformers.CLIPTextModel` the reference loads with `CLIPTextModel.from_pretrained(path,
subfolder="text_encoder")` (train.py:120): same parameter names (a Hugging Face checkpoint loads unchanged), same call
`model(input_ids)[0]` -> last_hidden_state (B, L, hidden).  Forward only: the text encoder is frozen on the finetune path
(train_text_encoder / use_text_lora are outside this build and rejected by train.main).

Per layer: LayerNorm -> fused Q|K|V GEMM (weights concatenated once: the encoder is frozen) -> causal attention (two batched
tcgen05 GEMMs around the row-softmax kernel with the causal mask; L = 77 makes this launch-bound, not worth a fused kernel)
-> output projection with the residual in the GEMM epilogue -> LayerNorm -> fc1 -> GELU -> fc2 (+ residual epilogue).
Embedding lookup (token + position) is one kernel; final LayerNorm as in CLIPTextTransformer."""
import json
import os
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import ops, prims


class _Attn(nn.Module):
    def __init__(self, C):
        super().__init__()
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = (nn.Linear(C, C) for _ in range(4))


class _Mlp(nn.Module):
    def __init__(self, C, I):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(C, I), nn.Linear(I, C)


class _Layer(nn.Module):
    def __init__(self, C, I, eps):
        super().__init__()
        self.self_attn = _Attn(C)
        self.layer_norm1 = nn.LayerNorm(C, eps=eps)
        self.mlp = _Mlp(C, I)
        self.layer_norm2 = nn.LayerNorm(C, eps=eps)


class _Embeddings(nn.Module):
    def __init__(self, vocab, positions, C):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, C)
        self.position_embedding = nn.Embedding(positions, C)


class _Encoder(nn.Module):
    def __init__(self, n, C, I, eps):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(C, I, eps) for _ in range(n)])


class _TextTransformer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.embeddings = _Embeddings(cfg.vocab_size, cfg.max_position_embeddings, cfg.hidden_size)
        self.encoder = _Encoder(cfg.num_hidden_layers, cfg.hidden_size, cfg.intermediate_size, cfg.layer_norm_eps)
        self.final_layer_norm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)


DEFAULTS = dict(vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=23, num_attention_heads=16,
                max_position_embeddings=77, hidden_act="gelu", layer_norm_eps=1e-5)   # the OpenCLIP ViT-H text tower of ms-1.7b


class CLIPTextModel(nn.Module):
    def __init__(self, config=None, **kwargs):
        super().__init__()
        cfg = dict(DEFAULTS)
        cfg.update(config if isinstance(config, dict) else (vars(config) if config is not None else {}))
        cfg.update(kwargs)
        self.config = SimpleNamespace(**{k: cfg[k] for k in DEFAULTS})
        if self.config.hidden_act not in ("gelu", "quick_gelu"):
            raise NotImplementedError(f"hidden_act {self.config.hidden_act!r}")
        if self.config.hidden_size % self.config.num_attention_heads or self.config.hidden_size % 8:
            raise ValueError("hidden_size must be a multiple of the head count and of 8")
        self.text_model = _TextTransformer(self.config)
        self.requires_grad_(False)
        self._packed = None

    @classmethod
    def from_pretrained(cls, path, subfolder=None, **unused):
        root = os.path.join(path, subfolder) if subfolder else path
        with open(os.path.join(root, "config.json")) as f:
            raw = json.load(f)
        model = cls({k: raw[k] for k in DEFAULTS if k in raw})
        st = os.path.join(root, "model.safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        else:
            sd = torch.load(os.path.join(root, "pytorch_model.bin"), map_location="cpu")
        sd = {k: v for k, v in sd.items() if not k.endswith("position_ids")}
        model.load_state_dict(sd)
        return model

    def load_state_dict(self, state_dict, strict=True):
        self._packed = None
        return super().load_state_dict(state_dict, strict=strict)

    @property
    def dtype(self):
        return torch.float32

    def _pack(self, device):
        """bf16 kernel-layout copies of the frozen weights, with q|k|v concatenated (one GEMM per layer instead of three)."""
        if self._packed is not None and self._packed["device"] == device:
            return self._packed
        bf = lambda w: prims.cast_f32_bf16(w.detach().float().contiguous()).view(w.shape[0], 1, 1, w.shape[1])  # noqa: E731
        layers = []
        for lyr in self.text_model.encoder.layers:
            a = lyr.self_attn
            layers.append(dict(
                qkv_w=bf(torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0)),
                qkv_b=torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0).detach().float().contiguous(),
                out_w=bf(a.out_proj.weight), out_b=a.out_proj.bias.detach().float().contiguous(),
                fc1_w=bf(lyr.mlp.fc1.weight), fc1_b=lyr.mlp.fc1.bias.detach().float().contiguous(),
                fc2_w=bf(lyr.mlp.fc2.weight), fc2_b=lyr.mlp.fc2.bias.detach().float().contiguous()))
        self._packed = dict(device=device, layers=layers)
        return self._packed

    @torch.no_grad()
    def forward(self, input_ids, attention_mask=None, **unused):
        """input_ids (B, L) int64 -> (last_hidden_state (B, L, hidden) fp32,).  The causal mask is always applied and, like
        the reference's call (train.py:786), no padding mask is."""
        cfg = self.config
        ids = input_ids.to(torch.int64).contiguous()
        B, L = ids.shape
        C, H = cfg.hidden_size, cfg.num_attention_heads
        D = C // H
        emb = self.text_model.embeddings
        pk = self._pack(ids.device)
        x = prims.embed_tokens(ids, emb.token_embedding.weight.detach().float().contiguous(),
                               emb.position_embedding.weight.detach().float().contiguous())          # [B*L, C] bf16
        ld = (L + 7) // 8 * 8
        for lyr, w in zip(self.text_model.encoder.layers, pk["layers"]):
            n, _ = prims.layernorm_fwd(x, lyr.layer_norm1.weight.detach().float(), lyr.layer_norm1.bias.detach().float(), lyr.layer_norm1.eps)
            qkv = prims.conv_fwd(n.view(1, 1, B * L, C), w["qkv_w"], w["qkv_b"]).view(B, L, 3 * C)
            q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
            s = torch.empty((B, H, L, ld), device=x.device, dtype=torch.float32)
            prims.bgemm(q, (1, q.stride(1), q.stride(0), D), k, (1, k.stride(1), k.stride(0), D), s, (ld, H * L * ld, L * ld),
                        L, L, D, B, H, D ** -0.5, 1)
            p = prims.softmax_fwd(s, L, ld, causal_period=L)
            a = torch.empty((B, L, C), device=x.device, dtype=x.dtype)
            prims.bgemm(p, (1, ld, H * L * ld, L * ld), v, (0, v.stride(1), v.stride(0), D), a, (C, L * C, D), L, D, L, B, H, 1.0, 0)
            x = prims.conv_fwd(a.view(1, 1, B * L, C), w["out_w"], w["out_b"], None, x.view(1, 1, B * L, C)).view(B * L, C)
            n, _ = prims.layernorm_fwd(x, lyr.layer_norm2.weight.detach().float(), lyr.layer_norm2.bias.detach().float(), lyr.layer_norm2.eps)
            h = prims.conv_fwd(n.view(1, 1, B * L, C), w["fc1_w"], w["fc1_b"]).view(B * L, -1)
            h = prims.gelu_bf16(h, quick=cfg.hidden_act == "quick_gelu")
            x = prims.conv_fwd(h.view(1, 1, B * L, h.shape[-1]), w["fc2_w"], w["fc2_b"], None, x.view(1, 1, B * L, C)).view(B * L, C)
        fl = self.text_model.final_layer_norm
        out, _ = prims.layernorm_fwd(x, fl.weight.detach().float(), fl.bias.detach().float(), fl.eps)
        return (out.float().view(B, L, C),)
