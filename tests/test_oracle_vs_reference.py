"""Pins the oracle's wiring restatement (oracle/unet3d_ref.py) against the reference's OWN models/*.py imported
unmodified (possible only where /root/reference exists; skipped on the GPU box)."""
import pytest
import torch

from oracle import unet3d_ref as R
from oracle.reference_import import import_reference_unet, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference sources not present on this machine")

SMALL = dict(block_out_channels=(64, 128, 128, 128), attention_head_dim=32, cross_attention_dim=64)


def _ref_model(cfg, seed=0):
    torch.manual_seed(seed)
    m = import_reference_unet()(**cfg).eval()
    for n, p in m.named_parameters():
        if ".conv4.3." in n:
            torch.nn.init.normal_(p, std=0.05)
    return m


@pytest.mark.parametrize("frames,hw", [(4, 16), (1, 8), (3, 12)])
def test_wiring_matches_reference(frames, hw):
    m = _ref_model(SMALL)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    torch.manual_seed(1)
    x = torch.randn(2, 4, frames, hw, hw)
    t = torch.tensor([500, 3])
    ehs = torch.randn(2, 7, 64)
    with torch.no_grad():
        y_ref = m(x, t, ehs).sample
        y = R.unet3d_forward(sd, R.full_config(**SMALL), x, t, ehs)
    assert y.shape == y_ref.shape
    assert (y - y_ref).abs().max().item() <= 2e-5 * y_ref.abs().max().item()


def test_checkpointed_reference_equals_plain():
    m = _ref_model(SMALL)
    torch.manual_seed(2)
    x, t, ehs = torch.randn(1, 4, 2, 8, 8), torch.tensor([10]), torch.randn(1, 7, 64)
    with torch.no_grad():
        a = m(x, t, ehs).sample
    m._set_gradient_checkpointing(True)
    b = m(x, t, ehs).sample
    assert torch.equal(a, b.detach())


def test_structural_pins_full_size():
    """1,411,233,860 parameters / 1,480 tensors for the ms-1.7b configuration, same keys+shapes as the product model."""
    from t2v_b200.models.unet_3d_condition import UNet3DConditionModel
    with torch.device("meta"):
        ref = import_reference_unet()()
        mine = UNet3DConditionModel()
    a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    assert len(a) == 1480 and sum(torch.Size(s).numel() for s in a.values()) == 1_411_233_860
    assert a == b
