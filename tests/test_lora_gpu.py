"""GPU parity of the LoRA training path: the cloneofsimo wrappers injected into the B200-native UNet run
their low-rank branch on the CUDA kernels (alpha epilogue, residual fusion); prediction and d loss / d lora must agree
with the SAME host wiring evaluated on CPU in fp32 over the per-primitive restatement (oracle/ops_ref.py), which
tests/test_lora_cpu.py in turn pins against the reference's own utils/lora.py.
Tolerances as in tests/test_unet_gpu.py (bf16 storage vs fp32): prediction rel-L2 < 4e-2, cosine > 0.999; >= 95% of
LoRA gradient tensors with cosine > 0.98."""
import contextlib
import io

import pytest
import torch

from helpers import cosine, emulated_prims, rel_l2, seeded_state_dict
from oracle import ops_ref

pytestmark = pytest.mark.gpu

SMALL = dict(block_out_channels=(64, 128, 128, 128), attention_head_dim=64, cross_attention_dim=64)


def _lora_model(device):
    from t2v_b200.models.unet_3d_condition import UNet3DConditionModel
    from t2v_b200.utils import lora as mylora
    m = UNet3DConditionModel(**SMALL)
    m.load_state_dict(seeded_state_dict(m, 0))
    with contextlib.redirect_stdout(io.StringIO()):
        params, _ = mylora.inject_trainable_lora_extended(m, mylora.UNET_EXTENDED_TARGET_REPLACE, r=8)
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for n, p in sorted(m.named_parameters()):
            if "lora_up" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            elif "lora_down" in n:
                p.copy_(torch.randn(p.shape, generator=g) / p[0].numel() ** 0.5)
    return m.to(device).eval()   # eval: the wrappers' dropout is off, everything else is unaffected (no BatchNorm)


def _run(m, x, t, ehs, target):
    pred = m(x, t, ehs).sample
    loss = torch.nn.functional.mse_loss(pred.float(), target.float())
    loss.backward()
    grads = {n: p.grad.detach().float().cpu() for n, p in m.named_parameters() if "lora" in n and p.grad is not None}
    return loss.item(), pred.detach().float().cpu(), grads


def test_lora_branch_prediction_and_gradients():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 4, 4, 16, 16, generator=g)
    t = torch.tensor([437])
    ehs = torch.randn(1, 7, 64, generator=g)
    target = torch.randn(1, 4, 4, 16, 16, generator=g)
    old = ops_ref.BF
    ops_ref.BF = torch.float32
    try:
        with emulated_prims():
            loss_r, pred_r, grads_r = _run(_lora_model("cpu"), x, t, ehs, target)
    finally:
        ops_ref.BF = old
    loss, pred, grads = _run(_lora_model("cuda"), x.cuda(), t.cuda(), ehs.cuda(), target.cuda())
    assert abs(loss - loss_r) <= 2e-3 * abs(loss_r), (loss, loss_r)
    assert rel_l2(pred, pred_r) < 4e-2 and cosine(pred, pred_r) > 0.999, (rel_l2(pred, pred_r), cosine(pred, pred_r))
    assert len(grads) == len(grads_r) and len(grads) > 100
    top = max(v.norm().item() for v in grads_r.values())
    cos = [cosine(grads[n], grads_r[n]) for n in grads_r if grads_r[n].norm().item() > 1e-4 * top]
    assert len(cos) > 100 and sum(c > 0.98 for c in cos) >= 0.95 * len(cos), sorted(cos)[:10]
