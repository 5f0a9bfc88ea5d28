"""LoRA path against the REFERENCE's own classes.

tests/golden/lora_*.pt were produced by /root/reference/utils/lora.py (LoraInjectedLinear / Conv2d / Conv3d and
inject_trainable_lora_extended, imported unmodified - tests/golden/make_golden_lora.py) on the reference's models/*.py, so
these cases pin the LoRA path to the reference implementation, not to this repo's own wiring (round-1 verdict).
CPU variants run the host wiring over the emulated primitives (fp32); GPU variants run the CUDA kernels (bf16 tolerances)."""
import contextlib
import io
import os

import pytest
import torch

from helpers import cosine, emulated_prims, rel_l2, seeded_state_dict
from oracle import ops_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEVICES = ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)]


@contextlib.contextmanager
def _backend(device):
    if device == "cpu":
        old = ops_ref.BF
        ops_ref.BF = torch.float32
        try:
            with emulated_prims():
                yield torch.float32
        finally:
            ops_ref.BF = old
    else:
        yield torch.bfloat16


def _close(a, b, tol, what):
    a, b = a.float().cpu(), b.float().cpu()
    err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)
    assert err < tol, f"{what}: rel-to-max error {err:.3e} >= {tol}"


@pytest.mark.parametrize("device", DEVICES)
@pytest.mark.parametrize("name", ["linear", "linear_nobias_r4", "conv2d", "conv2d_s2", "conv3d"])
def test_lora_wrappers_match_reference_classes(name, device):
    """y and the gradients w.r.t. x, lora_up, lora_down of one wrapped layer vs the reference wrapper's own autograd."""
    from t2v_b200 import ops
    from t2v_b200.utils import lora as mylora
    c = torch.load(os.path.join(GOLDEN, "lora_modules.pt"), weights_only=False)[name]
    st = c["state"]
    r, scale = c["r"], c["scale"]
    if name.startswith("linear"):
        w = st["linear.weight"]
        m = mylora.LoraInjectedLinear(w.shape[1], w.shape[0], "linear.bias" in st, r=r, dropout_p=0.1, scale=scale)
    elif name.startswith("conv2d"):
        w = st["conv.weight"]
        stride = 2 if name.endswith("s2") else 1
        m = mylora.LoraInjectedConv2d(w.shape[1], w.shape[0], 3, stride, 1, bias=True, r=r, dropout_p=0.1, scale=scale)
    else:
        w = st["conv.weight"]
        m = mylora.LoraInjectedConv3d(w.shape[1], w.shape[0], (3, 1, 1), (1, 0, 0), bias=True, r=r, dropout_p=0.1, scale=scale)
    m.load_state_dict(st)
    m = m.to(device).eval()
    for n, p in m.named_parameters():
        p.requires_grad_("lora" in n)      # the base layer is frozen, as after injection
    with _backend(device) as act:
        _check_wrapper(name, m, c, device, act)


def _check_wrapper(name, m, c, device, act):
    from t2v_b200.utils import lora as mylora
    tol = 1.5e-2 if device == "cuda" else 1e-4
    x = c["x"].to(device)
    # the wrappers run on channels-last bf16 activations inside the model; feed them the way layers.run_linear / run_conv do
    if name.startswith("linear"):
        xin = x.reshape(-1, x.shape[-1]).to(act).contiguous().requires_grad_(True)
        y = mylora.lora_linear_forward(m, xin)
        y_ref, dy = c["y"].reshape(-1, c["y"].shape[-1]), c["dy"].reshape(-1, c["dy"].shape[-1])
        dx_ref = c["dx"].reshape(-1, x.shape[-1])
        y.backward(dy.to(device).to(act))
        dx = xin.grad
    elif name.startswith("conv2d"):
        xin = x.permute(0, 2, 3, 1).to(act).contiguous().requires_grad_(True)
        y = mylora.lora_conv_forward(m, xin)
        y_ref, dx_ref = c["y"].permute(0, 2, 3, 1), c["dx"].permute(0, 2, 3, 1)
        y.backward(c["dy"].permute(0, 2, 3, 1).to(device).to(act).contiguous())
        dx = xin.grad
    else:   # (B, C, F, H, W) -> [B, F, H*W, C]
        B, C, F, H, W = x.shape
        to_cl = lambda t: t.permute(0, 2, 3, 4, 1).reshape(B, F, H * W, t.shape[1])  # noqa: E731
        xin = to_cl(x).to(act).contiguous().requires_grad_(True)
        y = mylora.lora_conv_forward(m, xin, pads=(1, 1, 0, 0))
        y_ref, dx_ref = to_cl(c["y"]), to_cl(c["dx"])
        y.backward(to_cl(c["dy"]).to(device).to(act).contiguous())
        dx = xin.grad
    _close(y, y_ref, tol, f"{name}: y")
    _close(dx, dx_ref, 1.4 * tol, f"{name}: dx")
    for n, g_ref in c["grads"].items():
        if "lora" not in n:
            continue
        g = dict(m.named_parameters())[n].grad
        _close(g.reshape(g_ref.shape), g_ref, 1.4 * tol, f"{name}: d {n}")


@pytest.mark.parametrize("device", DEVICES)
def test_lora_unet_matches_reference_injector_and_classes(device):
    """Whole-model LoRA: the reference UNet wiring + the reference injector + the reference wrapper classes (fixture) vs the
    B200-native UNet + this repo's injector on the GPU: loss, prediction, every LoRA gradient norm, 24 full gradients."""
    from t2v_b200 import step as S
    from t2v_b200.models.unet_3d_condition import UNet3DConditionModel
    from t2v_b200.utils import lora as mylora
    c = torch.load(os.path.join(GOLDEN, "lora_unet_small_f4.pt"), weights_only=False)
    m = UNet3DConditionModel(**c["cfg"])
    m.load_state_dict(seeded_state_dict(m, c["base_seed"]))
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
    m = m.to(device).eval()
    dev = device
    with _backend(device):
        loss, pred = S.finetune_loss(m, c["latents"].to(dev), c["noise"].to(dev), c["timesteps"].to(dev), c["text"].to(dev),
                                     S.ddpm_alphas_cumprod(device=dev), return_pred=True)
        loss.backward()
    # Loss tolerance 3e-3 on THIS model: its loss is a mean over only 4 x 4 x 16 x 16 prediction elements, so the bf16 error of the
    # prediction (rel-L2 2.3e-2) does not average out the way it does at the benchmark size, and the split-K reductions
    # (red.global.add order) make it vary from run to run: 5.3e-4 .. 1.3e-3 over six runs of tools/debug_lora_loss.py on one
    # B200.  The 1e-3 bar of north_star is asserted where it is meaningful: tests/test_parity_full_gpu.py (measured 4e-5 .. 1e-4).
    assert abs(loss.item() - c["loss"].item()) <= 3e-3 * abs(c["loss"].item()), (loss.item(), c["loss"].item())
    assert rel_l2(pred.float().cpu(), c["pred"]) < 4e-2 and cosine(pred.float().cpu(), c["pred"]) > 0.999
    params = dict(m.named_parameters())
    assert sum(1 for n, p in params.items() if "lora" in n and p.grad is not None) == c["n_lora"]
    top = max(c["grad_norms"].values())
    rel = [abs(params[n].grad.float().norm().item() - gn) / gn for n, gn in c["grad_norms"].items() if gn > 1e-3 * top]
    rel.sort()
    assert len(rel) > 100 and rel[len(rel) // 2] < 2e-2 and rel[int(0.95 * len(rel))] < 0.1, (len(rel), rel[len(rel) // 2], rel[-5:])
    for n, g_ref in c["grads"].items():
        assert cosine(params[n].grad.float().cpu(), g_ref) > 0.98, n
