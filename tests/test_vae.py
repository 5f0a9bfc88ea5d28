"""AutoencoderKL.encode / tensor_to_vae_latent: structural pin (34,163,664 parameters), CPU wiring
check with emulated primitives, and GPU parity against the oracle (oracle/leaves.py: vae_encode_moments)."""
import pytest
import torch

from helpers import cosine, emulated_prims, rel_l2, seeded_state_dict
from oracle import leaves as L
from oracle import ops_ref
from oracle import unet3d_ref as R

TINY = dict(block_out_channels=(32, 64, 64, 64), layers_per_block=1)


def _vae(cfg, seed=0):
    from t2v_b200.vae import AutoencoderKL
    m = AutoencoderKL(**cfg)
    sd = seeded_state_dict(m, seed)
    m.load_state_dict(sd)
    return m.eval(), sd


def test_parameter_census_matches_sd_vae():
    from t2v_b200.vae import AutoencoderKL
    with torch.device("meta"):
        m = AutoencoderKL()
    assert sum(p.numel() for p in m.parameters()) == 34_163_664


def test_cpu_wiring_exact():
    old, ops_ref.BF = ops_ref.BF, torch.float32
    try:
        from t2v_b200.vae import tensor_to_vae_latent
        m, sd = _vae(TINY)
        g = torch.Generator().manual_seed(1)
        pix = torch.rand(1, 3, 3, 32, 48, generator=g) * 2 - 1
        with emulated_prims():
            mom = m.encode_moments(pix.view(3, 3, 32, 48))
            lat = tensor_to_vae_latent(pix, m, generator=torch.Generator().manual_seed(5))
        ref = L.vae_encode_moments(sd, pix.view(3, 3, 32, 48), TINY["block_out_channels"], TINY["layers_per_block"])
        assert rel_l2(mom.float().permute(0, 3, 1, 2), ref) < 1e-5
        eps = torch.randn((1, 4, 3, 4, 6), generator=torch.Generator().manual_seed(5))
        assert rel_l2(lat, R.tensor_to_vae_latent(sd, pix, eps.permute(0, 2, 1, 3, 4).reshape(3, 4, 4, 6), TINY["block_out_channels"], 1)) < 1e-5
    finally:
        ops_ref.BF = old


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,n,hw", [(TINY, 4, (64, 64)), (dict(block_out_channels=(64, 128, 256, 256), layers_per_block=2), 2, (128, 96))])
def test_gpu_encode_matches_oracle(cfg, n, hw):
    m, sd = _vae(cfg, 2)
    m = m.cuda()
    g = torch.Generator().manual_seed(3)
    pix = torch.rand(n, 3, hw[0], hw[1], generator=g) * 2 - 1
    mom = m.encode_moments(pix.cuda()).float().cpu().permute(0, 3, 1, 2)
    ref = L.vae_encode_moments(sd, pix, cfg["block_out_channels"], cfg["layers_per_block"])
    assert rel_l2(mom, ref) < 4e-2 and cosine(mom, ref) > 0.999, (rel_l2(mom, ref), cosine(mom, ref))
    lat = m.encode(pix.cuda()).latent_dist.mode().cpu()
    assert rel_l2(lat, ref[:, :4]) < 4e-2
