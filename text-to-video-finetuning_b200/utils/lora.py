"""cloneofsimo-style LoRA for the B200-native UNet: same public names and file formats as the reference's
utils/lora.py (LoraInjectedLinear / LoraInjectedConv2d / LoraInjectedConv3d, inject_trainable_lora_extended,
extract_lora_ups_down, save_lora_weight, collapse_lora, monkeypatch_remove_lora, monkeypatch_or_replace_lora_extended,
tune_lora_scale, set_lora_diag ...), rebuilt so that the low-rank branch runs on the same sm_100a kernels as its host
layer:

    y = base(x) + dropout(up(selector(down(x)))) * scale          (reference utils/lora.py:57-62,134-139,211-216)

  * down / up are calls of the tcgen05 implicit-GEMM kernel (rank r rounds up to the 16-column UMMA minimum);
  * `* scale` and `+ base(x)` are folded into the up-projection's epilogue (alpha, residual) when dropout is inactive,
    and into one fused Philox dropout-scale-add kernel when it is active;
  * the injection contract is unchanged: wrappers are found by ancestor *class name*, only exact nn.Linear / nn.Conv2d /
    nn.Conv3d children are wrapped, base weight/bias Parameters are shared, lora_up starts at zero, lora_down ~ N(0, 1/r),
    rank is clamped to min(in, out), default dropout 0.1 (Linear, Conv2d) / 0 (Conv3d), scale 1.0.
"""
import json
import os
from typing import List, Optional, Set, Type, Union

import torch
import torch.nn as nn

from .. import ops

UNET_DEFAULT_TARGET_REPLACE = {"CrossAttention", "Attention", "GEGLU"}
UNET_EXTENDED_TARGET_REPLACE = {"ResnetBlock2D", "CrossAttention", "Attention", "GEGLU"}
TEXT_ENCODER_DEFAULT_TARGET_REPLACE = {"CLIPAttention"}
TEXT_ENCODER_EXTENDED_TARGET_REPLACE = {"CLIPAttention"}
DEFAULT_TARGET_REPLACE = UNET_DEFAULT_TARGET_REPLACE
EMBED_FLAG = "<embed>"


def _clamp_rank(r, a, b):
    lim = min(a, b)
    if r > lim:
        print(f"LoRA rank {r} is too large. setting to: {lim}")
        return lim
    return r


class _LoraWrapper(nn.Module):
    """Common state of the three wrappers: `r`, `dropout`, `scale`, `selector`, zero-initialised `lora_up`."""

    def _finish(self, r, dropout_p, scale):
        self.r = r
        self.dropout = nn.Dropout(dropout_p)
        self.scale = scale
        self.selector = nn.Identity()
        nn.init.normal_(self.lora_down.weight, std=1 / r)
        nn.init.zeros_(self.lora_up.weight)

    def realize_as_lora(self):
        return self.lora_up.weight.data * self.scale, self.lora_down.weight.data

    def _dropout_active(self):
        return self.training and self.dropout.p > 0


class LoraInjectedLinear(_LoraWrapper):
    def __init__(self, in_features, out_features, bias=False, r=4, dropout_p=0.1, scale=1.0):
        super().__init__()
        r = _clamp_rank(r, in_features, out_features)
        self.linear = nn.Linear(in_features, out_features, bias)
        self.lora_down = nn.Linear(in_features, r, bias=False)
        self.lora_up = nn.Linear(r, out_features, bias=False)
        self._finish(r, dropout_p, scale)

    def set_selector_from_diag(self, diag: torch.Tensor):
        assert diag.shape == (self.r,)
        self.selector = nn.Linear(self.r, self.r, bias=False)
        self.selector.weight.data = torch.diag(diag).to(self.lora_up.weight.device).to(self.lora_up.weight.dtype)

    def forward(self, input):
        return lora_linear_forward(self, input)


class LoraInjectedConv2d(_LoraWrapper):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True, r=4,
                 dropout_p=0.1, scale=1.0):
        super().__init__()
        r = _clamp_rank(r, in_channels, out_channels)
        mk = lambda cout, with_bias: nn.Conv2d(in_channels, cout, kernel_size, stride, padding, dilation, groups, with_bias)
        self.conv = mk(out_channels, bias)
        self.lora_down = ops_channels_last(mk(r, False))
        self.lora_up = ops_channels_last(nn.Conv2d(r, out_channels, 1, 1, 0, bias=False))
        self._finish(r, dropout_p, scale)

    def set_selector_from_diag(self, diag: torch.Tensor):
        assert diag.shape == (self.r,)
        self.selector = nn.Conv2d(self.r, self.r, 1, 1, 0, bias=False)
        self.selector.weight.data = torch.diag(diag).reshape(self.r, self.r, 1, 1).to(self.lora_up.weight.device).to(self.lora_up.weight.dtype)

    def forward(self, input):
        return lora_conv_forward(self, input)


class LoraInjectedConv3d(_LoraWrapper):
    def __init__(self, in_channels, out_channels, kernel_size=(3, 1, 1), padding=(1, 0, 0), bias=False, r=4, dropout_p=0,
                 scale=1.0):
        super().__init__()
        r = _clamp_rank(r, in_channels, out_channels)
        self.kernel_size, self.padding = kernel_size, padding
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size=kernel_size, padding=padding)
        self.lora_down = ops_channels_last(nn.Conv3d(in_channels, r, kernel_size=kernel_size, bias=False, padding=padding))
        self.lora_up = ops_channels_last(nn.Conv3d(r, out_channels, kernel_size=1, stride=1, padding=0, bias=False))
        self._finish(r, dropout_p, scale)

    def set_selector_from_diag(self, diag: torch.Tensor):
        assert diag.shape == (self.r,)
        self.selector = nn.Conv3d(self.r, self.r, 1, 1, 0, bias=False)
        self.selector.weight.data = torch.diag(diag).reshape(self.r, self.r, 1, 1, 1).to(self.lora_up.weight.device).to(self.lora_up.weight.dtype)

    def forward(self, input):
        return lora_conv_forward(self, input, pads=(1, 1, 0, 0))


def ops_channels_last(conv):
    from ..layers import _channels_last_
    return _channels_last_(conv)


_WRAPPERS = (LoraInjectedLinear, LoraInjectedConv2d, LoraInjectedConv3d)


# ------------------------------------------------------------------------------------------------ fused forward paths
def _finish_branch(m, u_fn, base):
    """y = base + scale * dropout(u).  u_fn(residual, alpha) runs the up-projection with that epilogue."""
    if not m._dropout_active():
        return u_fn(base, m.scale)
    return ops.dropout_scale_add(u_fn(None, 1.0), base, m.dropout.p, m.scale)


def lora_linear_forward(m, x, residual=None, out_fp32=False, stats_rows=0):
    if out_fp32:  # only the tiny per-clip time_emb_proj rows ask for fp32: widen the bf16 result (a [B, C] tensor)
        return lora_linear_forward(m, x, residual, False).float()
    x1, x2 = ops.fork(x)
    base = ops.linear(x1, m.linear.weight, m.linear.bias, residual)
    z = ops.linear(x2, m.lora_down.weight)
    if isinstance(m.selector, nn.Linear):
        z = ops.linear(z, m.selector.weight)
    # GroupNorm statistics ride on the last GEMM of the wrapper (the up-projection, whose epilogue adds the base output)
    return _finish_branch(m, lambda res, alpha: ops.linear(z, m.lora_up.weight, None, res, alpha=alpha,
                                                           stats_rows=stats_rows if res is not None else 0), base)


def lora_conv_forward(m, x, rowbias=None, residual=None, stride=None, pads=None, rb_div=1, cin_pad=0, cout_pad=0, stats_rows=0):
    conv = m.conv
    if stride is None:
        stride = conv.stride[0]
    if pads is None:
        p = conv.padding
        pads = (p[0], p[0], p[1], p[1]) if len(p) == 2 else (p[0], p[0], 0, 0)
    x1, x2 = ops.fork(x)
    base = ops.conv(x1, conv.weight, conv.bias, rowbias, residual, stride, pads, rb_div)
    z = ops.conv(x2, m.lora_down.weight, None, None, None, stride, pads)
    if not isinstance(m.selector, nn.Identity):
        z = ops.conv(z, m.selector.weight, pads=(0, 0, 0, 0))
    return _finish_branch(m, lambda res, alpha: ops.conv(z, m.lora_up.weight, None, None, res, 1, (0, 0, 0, 0), alpha=alpha,
                                                         stats_rows=stats_rows if res is not None else 0), base)


# ------------------------------------------------------------------------------------------------ module search
def _find_modules_v2(model, ancestor_class: Optional[Set[str]] = None, search_class: List[Type[nn.Module]] = [nn.Linear],
                     exclude_children_of: Optional[List[Type[nn.Module]]] = list(_WRAPPERS)):
    """Yield (parent, child_name, child) for every `search_class` instance below a module whose *class name* is in
    `ancestor_class` (all modules when None), skipping children of already-injected wrappers."""
    if ancestor_class is not None:
        ancestors = [mod for mod in model.modules() if mod.__class__.__name__ in ancestor_class]
    else:
        ancestors = list(model.modules())
    for anc in ancestors:
        for fullname, mod in anc.named_modules():
            if not isinstance(mod, tuple(search_class)):
                continue
            *path, name = fullname.split(".")
            parent = anc
            for part in path:
                parent = parent.get_submodule(part)
            if exclude_children_of and isinstance(parent, tuple(exclude_children_of)):
                continue
            yield parent, name, mod


_find_modules = _find_modules_v2


def _wrap(child, r):
    """Build the wrapper for an exact nn.Linear / nn.Conv2d / nn.Conv3d (subclasses are skipped, reference :458-462)."""
    cls = child.__class__
    if cls == nn.Linear:
        w = LoraInjectedLinear(child.in_features, child.out_features, child.bias is not None, r=r)
        w.linear.weight = child.weight
        if child.bias is not None:
            w.linear.bias = child.bias
    elif cls == nn.Conv2d:
        w = LoraInjectedConv2d(child.in_channels, child.out_channels, child.kernel_size, child.stride, child.padding,
                               child.dilation, child.groups, child.bias is not None, r=r)
        w.conv.weight = child.weight
        if child.bias is not None:
            w.conv.bias = child.bias
    elif cls == nn.Conv3d:
        w = LoraInjectedConv3d(child.in_channels, child.out_channels, bias=child.bias is not None,
                               kernel_size=child.kernel_size, padding=child.padding, r=r)
        w.conv.weight = child.weight
        if child.bias is not None:
            w.conv.bias = child.bias
    else:
        return None
    w.lora_down.to(child.weight.device, child.weight.dtype)
    w.lora_up.to(child.weight.device, child.weight.dtype)
    return w


def inject_trainable_lora_extended(model: nn.Module, target_replace_module: Set[str] = UNET_EXTENDED_TARGET_REPLACE, r: int = 4,
                                   loras=None):
    """Inject LoRA wrappers below every module named in `target_replace_module`; returns (param iterators, names)."""
    require_grad_params, names = [], []
    if loras is not None:
        loras = torch.load(loras)
    for parent, name, child in list(_find_modules(model, target_replace_module, search_class=[nn.Linear, nn.Conv2d, nn.Conv3d])):
        w = _wrap(child, r)
        if w is None:
            continue
        parent._modules[name] = w
        if loras is not None:
            w.lora_up.weight = nn.Parameter(_like(loras.pop(0), w.lora_up.weight))
            w.lora_down.weight = nn.Parameter(_like(loras.pop(0), w.lora_down.weight))
        w.lora_up.weight.requires_grad = True
        w.lora_down.weight.requires_grad = True
        require_grad_params.append(w.lora_up.parameters())
        require_grad_params.append(w.lora_down.parameters())
        names.append(name)
    return require_grad_params, names


def inject_trainable_lora(model: nn.Module, target_replace_module: Set[str] = DEFAULT_TARGET_REPLACE, r: int = 4, loras=None,
                          verbose: bool = False, dropout_p: float = 0.0, scale: float = 1.0):
    """Linear-only injector (reference utils/lora.py:336-390)."""
    require_grad_params, names = [], []
    if loras is not None:
        loras = torch.load(loras)
    for parent, name, child in list(_find_modules(model, target_replace_module, search_class=[nn.Linear])):
        w = LoraInjectedLinear(child.in_features, child.out_features, child.bias is not None, r=r, dropout_p=dropout_p, scale=scale)
        w.linear.weight = child.weight
        if child.bias is not None:
            w.linear.bias = child.bias
        w.to(child.weight.device).to(child.weight.dtype)
        parent._modules[name] = w
        if loras is not None:
            w.lora_up.weight = nn.Parameter(_like(loras.pop(0), w.lora_up.weight))
            w.lora_down.weight = nn.Parameter(_like(loras.pop(0), w.lora_down.weight))
        w.lora_up.weight.requires_grad = True
        w.lora_down.weight.requires_grad = True
        require_grad_params.append(w.lora_up.parameters())
        require_grad_params.append(w.lora_down.parameters())
        names.append(name)
    return require_grad_params, names


def _like(t, ref):
    """A loaded tensor placed on ref's device/dtype and in ref's memory format (conv weights stay channels-last)."""
    t = t.to(ref.device, ref.dtype).reshape(ref.shape)
    if ref.dim() == 4:
        return t.contiguous(memory_format=torch.channels_last)
    if ref.dim() == 5:
        return t.contiguous(memory_format=torch.channels_last_3d)
    return t.contiguous()


# ------------------------------------------------------------------------------------------------ extract / save
def extract_lora_ups_down(model, target_replace_module=DEFAULT_TARGET_REPLACE):
    loras = [(c.lora_up, c.lora_down) for _, _, c in _find_modules(model, target_replace_module, search_class=list(_WRAPPERS))]
    if not loras:
        raise ValueError("No lora injected.")
    return loras


def extract_lora_as_tensor(model, target_replace_module=DEFAULT_TARGET_REPLACE, as_fp16=True):
    loras = []
    for _, _, c in _find_modules(model, target_replace_module, search_class=list(_WRAPPERS)):
        up, down = c.realize_as_lora()
        loras.append((up.to(torch.float16), down.to(torch.float16)) if as_fp16 else (up, down))
    if not loras:
        raise ValueError("No lora injected.")
    return loras


def save_lora_weight(model, path="./lora.pt", target_replace_module=DEFAULT_TARGET_REPLACE):
    """cloneofsimo .pt format: flat python list [up_0, down_0, up_1, down_1, ...] of fp32 CPU tensors."""
    weights = []
    for up, down in extract_lora_ups_down(model, target_replace_module=target_replace_module):
        weights.append(up.weight.detach().to("cpu", torch.float32).contiguous())
        weights.append(down.weight.detach().to("cpu", torch.float32).contiguous())
    torch.save(weights, path)


def save_lora_as_json(model, path="./lora.json"):
    weights = []
    for up, down in extract_lora_ups_down(model):
        weights.append(up.weight.detach().cpu().numpy().tolist())
        weights.append(down.weight.detach().cpu().numpy().tolist())
    with open(path, "w") as f:
        json.dump(weights, f)


# ------------------------------------------------------------------------------------------------ merge / remove / load
def collapse_lora(model, alpha=1.0):
    """Fold up @ down into the base weights (valid when dropout is off; reference :781-815)."""
    targets = UNET_EXTENDED_TARGET_REPLACE | TEXT_ENCODER_EXTENDED_TARGET_REPLACE
    for _, name, c in _find_modules(model, targets, search_class=list(_WRAPPERS)):
        base = c.linear if isinstance(c, LoraInjectedLinear) else c.conv
        delta = c.lora_up.weight.data.flatten(start_dim=1) @ c.lora_down.weight.data.flatten(start_dim=1)
        new = base.weight.data + alpha * delta.reshape(base.weight.shape).to(base.weight.dtype)
        base.weight = nn.Parameter(_like(new, base.weight))


def monkeypatch_remove_lora(model):
    """Replace every wrapper by a plain layer that shares the base weight/bias."""
    for parent, name, c in list(_find_modules(model, search_class=list(_WRAPPERS))):
        src = c.linear if isinstance(c, LoraInjectedLinear) else c.conv
        if isinstance(src, nn.Linear):
            plain = nn.Linear(src.in_features, src.out_features, src.bias is not None)
        elif isinstance(src, nn.Conv2d):
            plain = nn.Conv2d(src.in_channels, src.out_channels, src.kernel_size, src.stride, src.padding, src.dilation,
                              src.groups, src.bias is not None)
        else:
            plain = nn.Conv3d(src.in_channels, src.out_channels, kernel_size=src.kernel_size, padding=src.padding,
                              bias=src.bias is not None)
        plain.weight = src.weight
        if src.bias is not None:
            plain.bias = src.bias
        parent._modules[name] = plain


def monkeypatch_or_replace_lora_extended(model, loras, target_replace_module=DEFAULT_TARGET_REPLACE, r: Union[int, List[int]] = 4):
    """Wrap (or re-use wrappers of) every Linear/Conv2d/Conv3d under the targets and load [up, down, ...] weights."""
    search = [nn.Linear, nn.Conv2d, nn.Conv3d] + list(_WRAPPERS)
    for parent, name, c in list(_find_modules(model, target_replace_module, search_class=search)):
        if c.__class__ in (nn.Linear, nn.Conv2d, nn.Conv3d) or isinstance(c, _WRAPPERS):
            src = c
            if isinstance(c, _WRAPPERS):
                src = c.linear if isinstance(c, LoraInjectedLinear) else c.conv
            if len(loras) and loras[0].dim() != (2 if isinstance(src, nn.Linear) else src.weight.dim()):
                continue
            rank = r.pop(0) if isinstance(r, list) else r
            w = _wrap(src, rank)
            if w is None:
                continue
            parent._modules[name] = w
            w.lora_up.weight = nn.Parameter(_like(loras.pop(0), w.lora_up.weight))
            w.lora_down.weight = nn.Parameter(_like(loras.pop(0), w.lora_down.weight))


def monkeypatch_or_replace_lora(model, loras, target_replace_module=DEFAULT_TARGET_REPLACE, r: Union[int, List[int]] = 4):
    """Linear-only loader (reference :818-859)."""
    for parent, name, c in list(_find_modules(model, target_replace_module, search_class=[nn.Linear, LoraInjectedLinear])):
        src = c.linear if isinstance(c, LoraInjectedLinear) else c
        if src.__class__ != nn.Linear:
            continue
        rank = r.pop(0) if isinstance(r, list) else r
        w = _wrap(src, rank)
        parent._modules[name] = w
        w.lora_up.weight = nn.Parameter(_like(loras.pop(0), w.lora_up.weight))
        w.lora_down.weight = nn.Parameter(_like(loras.pop(0), w.lora_down.weight))


def tune_lora_scale(model, alpha: float = 1.0):
    for m in model.modules():
        if isinstance(m, _WRAPPERS):
            m.scale = alpha


def set_lora_diag(model, diag: torch.Tensor):
    for m in model.modules():
        if isinstance(m, _WRAPPERS):
            m.set_selector_from_diag(diag)


def train_patch_pipe(pipe, patch_unet, patch_text):
    """After saving, the reference collapses and strips LoRA from the *copied* pipeline (utils/lora.py:1225-1235)."""
    if patch_unet:
        print("LoRA : Patching Unet")
        collapse_lora(pipe.unet)
        monkeypatch_remove_lora(pipe.unet)
    if patch_text:
        print("LoRA : Patching text encoder")
        collapse_lora(pipe.text_encoder)
        monkeypatch_remove_lora(pipe.text_encoder)


def inspect_lora(model):
    moved = {}
    for name, m in model.named_modules():
        if isinstance(m, _WRAPPERS):
            ups, downs = m.lora_up.weight.data.clone(), m.lora_down.weight.data.clone()
            dist = (ups.flatten(1) @ downs.flatten(1)).flatten().abs().mean().item()
            moved.setdefault(name, []).append(dist)
    return moved
