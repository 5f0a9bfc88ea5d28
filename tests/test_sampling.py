"""Validation preview: DPM-Solver++ sampler invariants, VAE decoder parity with the oracle restatement
(CPU exact wiring / GPU bf16), and the preview written as .mp4 from train.main."""
import contextlib
import os

import pytest
import torch

from helpers import cosine, emulated_prims, rel_l2, seeded_state_dict
from oracle import leaves as L
from oracle import ops_ref

TINY_VAE = dict(block_out_channels=(32, 64, 64, 64), layers_per_block=1)


@pytest.mark.parametrize("steps", [4, 10, 25])
def test_dpm_solver_is_exact_for_a_point_mass(steps):
    """If the data distribution is a single point x*, eps(x, t) = (x - alpha_t x*) / sigma_t and every DPM-Solver++ step (first
    or second order) must keep x_t = alpha_t x* + sigma_t z with the SAME z - an end-to-end check of the update formulas."""
    from t2v_b200.sampling import DPMSolverMultistep
    ac = L.ddpm_alphas_cumprod()
    s = DPMSolverMultistep(ac, steps)
    g = torch.Generator().manual_seed(0)
    xstar, z = torch.randn(5, generator=g).double(), torch.randn(5, generator=g).double()
    t0 = int(s.timesteps[0])
    x = s.alpha[t0] * xstar + s.sigma[t0] * z
    for i, t in enumerate(s.timesteps.tolist()):
        eps = (x - s.alpha[t] * xstar) / s.sigma[t]
        x = s.step(eps, x)
        nxt = int(s.timesteps[i + 1]) if i + 1 < steps else 0
        want = s.alpha[nxt] * xstar + s.sigma[nxt] * z
        assert torch.allclose(x, want, rtol=1e-9, atol=1e-9), (i, (x - want).abs().max())
    assert (x - xstar).abs().max() < 0.05          # t = 0: alpha_0 ~ 0.9996, sigma_0 ~ 0.029


def _vae(device):
    from t2v_b200.vae import AutoencoderKL
    m = AutoencoderKL(**TINY_VAE, build_decoder=True)
    sd = seeded_state_dict(m, 4)
    m.load_state_dict(sd)
    return m.to(device).eval(), sd


@pytest.mark.parametrize("device", ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_vae_decoder_matches_oracle(device):
    old = ops_ref.BF
    ctx = emulated_prims() if device == "cpu" else contextlib.nullcontext()
    if device == "cpu":
        ops_ref.BF = torch.float32
    try:
        with ctx:
            m, sd = _vae(device)
            z = torch.randn(3, 4, 8, 12, generator=torch.Generator().manual_seed(2))
            img = m.decode(z.to(device)).sample.float().cpu()
    finally:
        ops_ref.BF = old
    ref = L.vae_decode(sd, z, TINY_VAE["block_out_channels"], TINY_VAE["layers_per_block"])
    assert img.shape == ref.shape == (3, 3, 64, 96)
    if device == "cpu":
        assert rel_l2(img, ref) < 1e-5, rel_l2(img, ref)
    else:
        assert rel_l2(img, ref) < 4e-2 and cosine(img, ref) > 0.999, (rel_l2(img, ref), cosine(img, ref))


def test_full_checkpoint_builds_the_decoder_and_encoder_only_stays_small():
    from t2v_b200.vae import AutoencoderKL
    full, sd = _vae("cpu")
    m = AutoencoderKL(**TINY_VAE)
    assert m.decoder is None
    m.load_state_dict(sd)                       # decoder keys present -> decoder appears
    assert m.decoder is not None and sum(p.numel() for p in m.parameters()) == sum(p.numel() for p in full.parameters())
    enc_only = {k: v for k, v in sd.items() if not k.startswith(("decoder.", "post_quant_conv."))}
    m2 = AutoencoderKL(**TINY_VAE)
    m2.load_state_dict(enc_only)
    assert m2.decoder is None


def test_validation_preview_from_train_main(tmp_path):
    """validation_data + validation_steps: train.main writes samples/<step>_<prompt>.mp4 (two sampler steps, tiny models)."""
    import test_pipeline_train as T
    from t2v_b200.vae import AutoencoderKL
    root = T._pipeline_folder(str(tmp_path / "pipe"))
    torch.manual_seed(1)
    AutoencoderKL(block_out_channels=(32, 32, 64, 64), layers_per_block=1, build_decoder=True).save_pretrained(os.path.join(root, "vae"))
    from t2v_b200 import train
    out = str(tmp_path / "out")
    with emulated_prims():
        train.main(pretrained_model_path=root, output_dir=out, dataset_types=["synthetic"],
                   train_data=dict(n=2, n_sample_frames=2, height=64, width=64), max_train_steps=1, learning_rate=1e-4,
                   checkpointing_steps=10, seed=0, shuffle=False, device="cpu", eval_train=True, trainable_modules=["attn1"],
                   load_side_models=True, validation_steps=1,
                   validation_data=dict(prompt="a dog", sample_preview=True, num_frames=2, width=32, height=32, num_inference_steps=2,
                                        guidance_scale=2.0))
    files = os.listdir(os.path.join(out, "samples"))
    assert len(files) == 1 and files[0].endswith(".mp4") and os.path.getsize(os.path.join(out, "samples", files[0])) > 0
