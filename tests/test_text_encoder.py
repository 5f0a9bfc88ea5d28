"""Frozen CLIP text encoder against the REAL `transformers.CLIPTextModel` (installed in this image: a
genuine third-party pin, not a restatement): same state dict, same token ids -> last_hidden_state.
CPU: host wiring over the emulated primitives (fp32, tight).  GPU: the CUDA kernels (bf16 tolerance)."""
import contextlib

import pytest
import torch

from helpers import cosine, emulated_prims, rel_l2
from oracle import ops_ref

CFGS = {"tiny": dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=1, vocab_size=100,
                     max_position_embeddings=16, hidden_act="gelu"),
        "vit_h_2layers": dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=2, num_attention_heads=16, vocab_size=1000,
                              max_position_embeddings=77, hidden_act="gelu"),
        "quick_gelu": dict(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2, vocab_size=50,
                           max_position_embeddings=24, hidden_act="quick_gelu")}


@contextlib.contextmanager
def _backend(device):
    if device == "cpu":
        old = ops_ref.BF
        ops_ref.BF = torch.float32
        try:
            with emulated_prims():
                yield
        finally:
            ops_ref.BF = old
    else:
        yield


@pytest.mark.parametrize("device", ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)])
@pytest.mark.parametrize("name", list(CFGS))
def test_matches_transformers_clip_text_model(name, device):
    from transformers import CLIPTextConfig
    from transformers import CLIPTextModel as HF
    from t2v_b200.text_encoder import CLIPTextModel
    cfg = CFGS[name]
    torch.manual_seed(0)
    hf = HF(CLIPTextConfig(**cfg)).eval()
    with torch.no_grad():   # default init is tiny (std 0.02): give the residual stream and the norms something to do
        for n, p in hf.named_parameters():
            if p.dim() >= 2 and "embedding" not in n:
                p.mul_(4.0)
            elif "layer_norm" in n and n.endswith("weight"):
                p.add_(0.2 * torch.randn_like(p))
    ids = torch.randint(0, cfg["vocab_size"], (3, cfg["max_position_embeddings"]))
    with torch.no_grad():
        want = hf(ids)[0]
    mine = CLIPTextModel(cfg)
    sd = {k: v for k, v in hf.state_dict().items() if not k.endswith("position_ids")}
    mine.load_state_dict(sd)
    mine = mine.to(device)
    with _backend(device):
        got = mine(ids.to(device))[0].float().cpu()
    assert got.shape == want.shape
    if device == "cpu":
        assert rel_l2(got, want) < 1e-5, rel_l2(got, want)
    else:
        assert rel_l2(got, want) < 3e-2 and cosine(got, want) > 0.999, (rel_l2(got, want), cosine(got, want))


def test_from_pretrained_reads_a_hugging_face_folder(tmp_path):
    from transformers import CLIPTextConfig
    from transformers import CLIPTextModel as HF
    from t2v_b200.text_encoder import CLIPTextModel
    hf = HF(CLIPTextConfig(**CFGS["tiny"]))
    hf.save_pretrained(str(tmp_path / "m" / "text_encoder"))
    mine = CLIPTextModel.from_pretrained(str(tmp_path / "m"), subfolder="text_encoder")
    for k, v in hf.state_dict().items():
        if not k.endswith("position_ids"):
            assert torch.equal(mine.state_dict()[k], v), k
    assert not any(p.requires_grad for p in mine.parameters())
