"""Fused AdamW on the flat arena (optim.FusedAdamW): the update rule (restated in oracle/ops_ref.adamw_step, which the CUDA
kernel csrc/optim.cu mirrors line by line) against torch.optim.AdamW, with several parameter groups, frozen parameters
and folded gradient clipping.  CPU always; the GPU variant is opt-in (T2V_TEST_OPTIN=1) until the kernel has run once."""
import pytest
import torch

from helpers import emulated_prims



def _net(device):
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(24, 40), torch.nn.Linear(40, 16, bias=False), torch.nn.LayerNorm(16), torch.nn.Linear(16, 8))
    net[1].weight.requires_grad_(False)          # a frozen matrix in the middle of the arena
    return net.to(device)


def _compare(device, rtol):
    from t2v_b200 import optim
    from t2v_b200.runtime import ParamArena
    ref, fused = _net(device), _net(device)
    fused.load_state_dict(ref.state_dict())
    arena = ParamArena(fused)
    groups = lambda m: [dict(params=[p for n, p in m.named_parameters() if n.startswith("0.")], lr=3e-3),  # noqa: E731
                        dict(params=[p for n, p in m.named_parameters() if not n.startswith("0.")], lr=1e-3, weight_decay=0.05)]
    o_ref = torch.optim.AdamW(groups(ref), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-2)
    o_fus = optim.FusedAdamW(arena, groups(fused), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-2, max_grad_norm=1.0)
    assert len(o_fus._sets) == 2 and o_fus.covers_all_trainable()
    g = torch.Generator().manual_seed(1)
    frozen_before = fused[1].weight.detach().clone()
    for step in range(4):
        arena.reattach_grads()
        arena.zero_grads()
        for (n, p), (_, q) in zip(ref.named_parameters(), fused.named_parameters()):
            if not p.requires_grad:
                continue
            grad = torch.randn(p.shape, generator=g).to(device) * (3.0 if step == 2 else 0.3)
            p.grad = grad.clone()
            q.grad.copy_(grad)
        norm = torch.nn.utils.clip_grad_norm_([p for p in ref.parameters() if p.grad is not None], 1.0)
        o_ref.step()
        o_fus.step(zero_grad=True)     # clip factor, step count and bias corrections are computed on the device
        assert abs(float(o_fus.last_grad_norm()) - float(norm)) < 1e-5 * float(norm) and o_fus.steps == step + 1
        assert float(arena.grad.abs().max()) == 0.0 or all(float(q.grad.abs().max()) == 0.0 for q in fused.parameters() if q.requires_grad)
    for (n, p), (_, q) in zip(ref.named_parameters(), fused.named_parameters()):
        assert torch.allclose(p, q, rtol=rtol, atol=1e-7), (n, float((p - q).abs().max()))
    assert torch.equal(fused[1].weight, frozen_before)                       # frozen: untouched (no weight decay either)
    # the bf16 shadow of every trainable matrix follows the master weights
    for p in fused.parameters():
        if p.dim() >= 2 and p.requires_grad:
            assert torch.equal(p._t2v_shadow.reshape(-1).float(), p.detach().reshape(-1).bfloat16().float())


def test_fused_adamw_matches_torch_cpu():
    with emulated_prims():
        _compare("cpu", 1e-5)


@pytest.mark.gpu
def test_fused_adamw_matches_torch_gpu():
    _compare("cuda", 1e-5)
