"""Training-step glue of the hot path: the B200-native equivalents of the reference's
`tensor_to_vae_latent` (train.py:339-347), `sample_noise` (:349-358), `noise_scheduler.add_noise` (:760) and the
epsilon-MSE of `finetune_unet` (:720-836), plus the data-parallel step object used by train.py and bench.py."""
import os

import torch
import torch.distributed as dist

from . import ops, prims
from .runtime import GradientBuckets, GraphedStep, ParamArena, allreduce_gradients


def ddpm_alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, device=None):
    """'scaled_linear' DDPM schedule of the ms-1.7b / zeroscope scheduler config (DDPMScheduler.alphas_cumprod)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    a = torch.cumprod(1.0 - betas, dim=0)
    return a.to(device) if device is not None else a


def sample_noise(latents, noise_strength=0.0, use_offset_noise=False, generator=None):
    """train.py:349-358: Gaussian noise, optionally plus `strength * randn(B, C, F, 1, 1)` offset noise."""
    noise = torch.randn(latents.shape, device=latents.device, dtype=latents.dtype, generator=generator)
    if use_offset_noise:
        b, c, f = latents.shape[:3]
        noise = noise + noise_strength * torch.randn(b, c, f, 1, 1, device=latents.device, dtype=latents.dtype, generator=generator)
    return noise


def finetune_loss(unet, latents, noise, timesteps, encoder_hidden_states, alphas_cumprod, return_pred=False):
    """One UNet pass of finetune_unet for prediction_type 'epsilon':
       noisy = add_noise(latents, noise, t)  ->  pred = unet(noisy, t, text)  ->  mse(pred.float(), noise.float()).
    add_noise is fused into the layout-conversion kernel at the input, the loss reads the channels-last prediction
    directly, so no (B,C,F,H,W) activation is ever materialised."""
    B, C, F, H, W = latents.shape
    x = prims.latents_to_nhwc8(latents.float().contiguous(), noise.float().contiguous(), alphas_cumprod, timesteps.to(torch.int64).contiguous())
    text = unet.prepare_text(encoder_hidden_states)
    pred = unet.forward_channels_last(x, timesteps.to(torch.int64).contiguous(), text, B, F)
    loss = ops.mse_loss_nhwc8(pred, noise.float().contiguous())
    if return_pred:
        return loss, prims.nhwc8_to_latents(pred.detach(), B, C, F)
    return loss


class DataParallelStep:
    """fwd + bwd of `passes` UNet passes over one clip batch per rank, then ONE gradient all-reduce.

    `passes=2` reproduces the reference's two-pass video step (train.py:814-834, H3: loss = loss_0 + loss_1);
    throughput is reported per pass with passes=1."""

    def __init__(self, unet, alphas_cumprod, passes=1, use_graph=False, adopt=True):
        self.unet = unet
        self.abar = alphas_cumprod
        self.passes = passes
        self.arena = ParamArena(unet) if adopt else None
        self.use_graph = use_graph
        self.sync_gradients = True   # set False to run fwd+bwd only (profiling on a single rank)
        self._graph = None
        # world > 1: per-block gradient all-reduces are issued from inside the backward pass (overlap), see GradientBuckets
        self.buckets = None
        if (self.arena is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
                and not os.environ.get("T2V_NO_OVERLAP")):
            self.buckets = GradientBuckets(self.arena, unet)
            self.buckets.install()

    def _fwd_bwd(self, latents, noise, timesteps, text):
        if self.arena is not None:
            self.arena.zero_grads()
            self.arena.refresh_shadow()
        total = None
        overlap = self.buckets is not None and self.sync_gradients
        for i in range(self.passes):
            loss = finetune_loss(self.unet, latents, noise, timesteps, text, self.abar)
            if overlap:
                self.buckets.armed = i == self.passes - 1   # gradients are final only in the last pass
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        if overlap:
            self.buckets.finish()
        return total

    def __call__(self, latents, noise, timesteps, encoder_hidden_states):
        if self.arena is not None:
            self.arena.reattach_grads()
        args = (latents, noise, timesteps, encoder_hidden_states)
        if self.use_graph:
            if self._graph is None:
                self._graph = GraphedStep(self._fwd_bwd, args)
            loss = self._graph(*args)
        else:
            loss = self._fwd_bwd(*args)
        if self.arena is not None and self.sync_gradients and self.buckets is None:
            allreduce_gradients(self.arena)
        return loss
