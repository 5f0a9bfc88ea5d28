"""The product has no CPU or library fallback: without the CUDA path the model refuses to run, and no product module
imports the oracle (which is test infrastructure)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "text-to-video-finetuning_b200")


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    offenders = []
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                if pat.search(src):
                    offenders.append(os.path.relpath(os.path.join(base, f), ROOT))
    assert not offenders, offenders


def test_cpu_tensors_are_rejected_not_emulated():
    from t2v_b200.models.unet_3d_condition import UNet3DConditionModel
    m = UNet3DConditionModel(block_out_channels=(32, 64, 64, 64), attention_head_dim=32, cross_attention_dim=32).eval()
    x, t, ehs = torch.randn(1, 4, 2, 8, 8), torch.tensor([10]), torch.randn(1, 3, 32)
    with pytest.raises((AssertionError, RuntimeError)):
        m(x, t, ehs)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from t2v_b200 import native
    monkeypatch.setattr(native, "_lib", None, raising=False)
    monkeypatch.setattr(native, "LIB_PATH", str(tmp_path / "nope.so"), raising=False)
    with pytest.raises(Exception):
        native.lib()
