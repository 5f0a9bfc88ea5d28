#!/usr/bin/env python
"""Repeats the forward of tests/test_lora_reference_golden.py::test_lora_unet_matches_reference_injector_and_classes and
prints the loss next to the fixture's: run-to-run spread of the bf16 path on the small LoRA UNet."""
import contextlib
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import test_lora_reference_golden as T  # noqa: E402
from t2v_b200 import step as S  # noqa: E402
from t2v_b200.models.unet_3d_condition import UNet3DConditionModel  # noqa: E402
from t2v_b200.utils import lora as mylora  # noqa: E402

c = torch.load(os.path.join(T.GOLDEN, "lora_unet_small_f4.pt"), weights_only=False)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    m = UNet3DConditionModel(**c["cfg"])
    m.load_state_dict(T.seeded_state_dict(m, c["base_seed"]))
    m.requires_grad_(False)
    with contextlib.redirect_stdout(io.StringIO()):
        mylora.inject_trainable_lora_extended(m, {"UNet3DConditionModel"}, r=c["r"])
    g = torch.Generator().manual_seed(c["lora_seed"])
    with torch.no_grad():
        for n, p in sorted(m.named_parameters()):
            if "lora_up" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            elif "lora_down" in n:
                p.copy_(torch.randn(p.shape, generator=g) / p[0].numel() ** 0.5)
    m = m.to("cuda").eval()
    loss, pred = S.finetune_loss(m, c["latents"].cuda(), c["noise"].cuda(), c["timesteps"].cuda(), c["text"].cuda(),
                                 S.ddpm_alphas_cumprod(device="cuda"), return_pred=True)
    ref = c["loss"].item()
    print(f"rep {rep}: loss {loss.item():.7f}  golden {ref:.7f}  rel {abs(loss.item() - ref) / ref:.2e}  pred rel-L2 {T.rel_l2(pred.float().cpu(), c['pred']):.3e}",
          flush=True)
