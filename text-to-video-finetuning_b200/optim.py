"""Fused AdamW on the flat parameter arena (the reference builds torch.optim.AdamW at train.py:616-623).

One kernel call per contiguous run of trainable parameters that share hyper-parameters: reads p, g, m, v, writes p, m, v and
the bf16 compute shadow (so runtime.ParamArena.refresh_shadow becomes unnecessary for the ranges it covers).  Frozen
parameters are never touched (torch semantics: grad None => skipped).  Gradient clipping can be folded in through
`grad_scale` (see `clip_scale`).  Opt-in (train.main(fused_adamw=True)): the update rule is checked on CPU against
torch.optim.AdamW; the CUDA kernel has not been exercised on a GPU yet."""
import torch

from . import prims
from .runtime import _align


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, arena, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.arena = arena
        self.exp_avg = torch.zeros_like(arena.master)
        self.exp_avg_sq = torch.zeros_like(arena.master)
        self.steps = 0
        self._off = {id(p): o for p, o in zip(arena.params, arena.offsets)}
        for group in self.param_groups:
            for p in group["params"]:
                if id(p) not in self._off:
                    raise ValueError("FusedAdamW only drives parameters that live in the arena")

    def _runs(self, group):
        """Contiguous [a, b) ranges of this group's trainable parameters in arena order (alignment gaps between adjacent
        parameters hold zeros in master / grad / state and may be swept along)."""
        spans = sorted((self._off[id(p)], self._off[id(p)] + _align(p.numel())) for p in group["params"] if p.requires_grad)
        runs = []
        for a, b in spans:
            if runs and runs[-1][1] == a:
                runs[-1][1] = b
            else:
                runs.append([a, b])
        return runs

    def clip_scale(self, max_norm):
        """Factor that torch.nn.utils.clip_grad_norm_ would apply, without touching the gradients (pass it to step())."""
        total = torch.zeros((), device=self.arena.grad.device)
        for group in self.param_groups:
            for a, b in self._runs(group):
                total = total + self.arena.grad[a:b].double().pow(2).sum().float()
        return float(torch.clamp(max_norm / (total.sqrt() + 1e-6), max=1.0))

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0, zero_grad=False):
        loss = closure() if closure is not None else None
        self.steps += 1
        ar = self.arena
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for a, b in self._runs(group):
                hi = min(b, ar.n_mat)   # only matrices have a bf16 shadow (they come first in the arena)
                for lo_, hi_, sh in ((a, max(a, hi), True), (max(a, hi), b, False)):
                    if hi_ <= lo_:
                        continue
                    prims.adamw_step(ar.master[lo_:hi_], ar.grad[lo_:hi_], self.exp_avg[lo_:hi_], self.exp_avg_sq[lo_:hi_],
                                     ar.shadow[lo_:hi_] if sh else None, group["lr"], b1, b2, group["eps"], group["weight_decay"],
                                     self.steps, grad_scale, zero_grad)
        return loss
