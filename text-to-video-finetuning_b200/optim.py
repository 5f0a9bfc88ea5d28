"""Fused AdamW + global-norm gradient clipping on the flat parameter arena.

Reference: `torch.optim.AdamW` built at train.py:616-623, `accelerator.clip_grad_norm_(..., max_grad_norm)` at :868-876,
`optimizer.zero_grad()` at :879.  Here one optimizer step is three kernel launches per hyper-parameter set (csrc/optim.cu):
sum of squared gradients -> device-side scalars (step count, bias corrections, clip factor) -> update, which also writes the
bf16 compute shadow of every updated matrix (runtime.ParamArena.refresh_shadow is no longer needed per step) and zeroes the
gradient ranges it consumed (no per-step memset).  Nothing in `launch()` touches the host, so the optimizer is captured
into the CUDA graph of the training step; the learning rate reaches the device through `push_hyperparams()`.

Frozen parameters are never touched (torch semantics: grad None => skipped, no weight decay either).
"""
import torch

from . import prims
from .runtime import _align

CHUNK = 1 << 16   # elements per chunk-table entry


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, arena, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.arena = arena
        self.max_grad_norm = max_grad_norm
        dev = arena.master.device
        self.exp_avg = torch.zeros_like(arena.master)
        self.exp_avg_sq = torch.zeros_like(arena.master)
        self.state_dev = torch.zeros(1, device=dev, dtype=torch.int64)     # [0] optimizer step count
        self.sq = torch.zeros(2, device=dev, dtype=torch.float64)          # [0] sum g^2 (running), [1] last gradient norm
        self._off = {id(p): o for p, o in zip(arena.params, arena.offsets)}
        for group in self.param_groups:
            for p in group["params"]:
                if id(p) not in self._off:
                    raise ValueError("FusedAdamW only drives parameters that live in the arena")
        self._sets = None
        self._build()

    # ------------------------------------------------------------------------------------------------ chunk tables
    @staticmethod
    def _key(group):
        return (float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]), float(group["weight_decay"]))

    def _runs(self, params):
        """Contiguous [a, b) ranges of the trainable parameters among `params` in arena order (alignment gaps between
        adjacent parameters hold zeros in master / grad / state and are swept along)."""
        spans = sorted((self._off[id(p)], self._off[id(p)] + _align(p.numel())) for p in params if p.requires_grad)
        runs = []
        for a, b in spans:
            if runs and runs[-1][1] == a:
                runs[-1][1] = b
            else:
                runs.append([a, b])
        return runs

    def _build(self):
        """Groups with identical hyper-parameters form one set: one chunk table, one row of device hyper-parameters."""
        by_key = {}
        for gi, group in enumerate(self.param_groups):
            by_key.setdefault(self._key(group), []).append(gi)
        dev = self.arena.master.device
        n_mat = self.arena.n_mat
        sets = []
        for key, gis in by_key.items():
            params = [p for gi in gis for p in self.param_groups[gi]["params"]]
            chunks = []
            for a, b in self._runs(params):
                for lo, hi in ((a, min(b, n_mat)), (max(a, n_mat), b)):     # never straddle the matrix / vector boundary
                    pos = lo
                    while pos < hi:
                        n = min(CHUNK, hi - pos)
                        chunks.append((pos, n))
                        pos += n
            if not chunks:
                continue
            sets.append({"groups": gis, "key": key, "chunks": torch.tensor(chunks, dtype=torch.int64, device=dev).contiguous(),
                         "n": sum(n for _, n in chunks)})
        if not sets:
            raise ValueError("FusedAdamW: no trainable parameters")
        self._sets = sets
        self.hp_host = torch.zeros((len(sets), 5), dtype=torch.float32)
        if dev.type == "cuda":
            self.hp_host = self.hp_host.pin_memory()
        self.hp_in = torch.zeros((len(sets), 5), device=dev, dtype=torch.float32)
        self.hp = torch.zeros((len(sets), 8), device=dev, dtype=torch.float32)
        self.generation = getattr(self, "generation", 0) + 1   # graphs captured against an older table must be re-captured

    def covers_all_trainable(self):
        """True when every trainable parameter of the arena is updated (and therefore zeroed) by this optimizer."""
        mine = {id(p) for g in self.param_groups for p in g["params"]}
        return all(id(p) in mine for p in self.arena.params if p.requires_grad)

    @property
    def trainable_elements(self):
        return sum(s["n"] for s in self._sets)

    # ------------------------------------------------------------------------------------------------ step
    def push_hyperparams(self):
        """Host -> device copy of (lr, beta1, beta2, eps, weight_decay) per set (stream-ordered, before the step's kernels or
        the graph replay).  A scheduler that makes groups of one set diverge triggers a rebuild of the tables."""
        for s in self._sets:
            if any(self._key(self.param_groups[gi])[1:] != s["key"][1:] or
                   self.param_groups[gi]["lr"] != self.param_groups[s["groups"][0]]["lr"] for gi in s["groups"]):
                self._build()
                break
        for i, s in enumerate(self._sets):
            g = self.param_groups[s["groups"][0]]
            self.hp_host[i, 0] = float(g["lr"])
            self.hp_host[i, 1], self.hp_host[i, 2] = float(g["betas"][0]), float(g["betas"][1])
            self.hp_host[i, 3], self.hp_host[i, 4] = float(g["eps"]), float(g["weight_decay"])
        self.hp_in.copy_(self.hp_host, non_blocking=True)

    def launch(self, zero_grad=True, grad_bf16=None):
        """The device-only part of one step (capturable): gradient norm, scalars, update.  grad_bf16: flat bf16 twin of the
        gradient buffer holding the all-reduced gradients (runtime.GradientBuckets.comm); the fp32 buffer is then only zeroed."""
        ar = self.arena
        if self.max_grad_norm is not None:
            for s in self._sets:
                prims.sqnorm_chunks(ar.grad, s["chunks"], self.sq, grad_bf16)
        prims.adamw_prepare(self.hp_in, self.hp, self.state_dev, self.sq, self.max_grad_norm or 0.0)
        for i, s in enumerate(self._sets):
            prims.adamw_chunks(ar.master, ar.grad, self.exp_avg, self.exp_avg_sq, ar.shadow, ar.n_mat, s["chunks"], self.hp[i], zero_grad,
                               grad_bf16)

    @torch.no_grad()
    def step(self, closure=None, zero_grad=True):
        loss = closure() if closure is not None else None
        self.push_hyperparams()
        self.launch(zero_grad)
        return loss

    def zero_grad(self, set_to_none=False):
        """Gradients live in the arena and are zeroed by the update kernel; an explicit call zeroes the whole flat buffer."""
        self.arena.zero_grads()

    @property
    def steps(self):
        return int(self.state_dev.item())

    def last_grad_norm(self):
        """Global gradient norm seen by the last step (device scalar, fp64); only maintained when clipping is on."""
        return self.sq[1]

    # ------------------------------------------------------------------------------------------------ checkpointing
    def state_dict(self):
        d = super().state_dict()
        d["fused"] = {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "step": self.steps}
        return d

    def load_state_dict(self, state_dict):
        fused = state_dict.pop("fused", None)
        super().load_state_dict(state_dict)
        if fused is not None:
            self.exp_avg.copy_(fused["exp_avg"])
            self.exp_avg_sq.copy_(fused["exp_avg_sq"])
            self.state_dev.fill_(int(fused["step"]))
