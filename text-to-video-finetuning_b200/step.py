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
    """One optimisation micro-step per call: fwd + bwd of `passes` UNet passes over one clip batch per rank; on the boundary
    of a gradient-accumulation window ONE gradient all-reduce and, when an optimizer is attached, the fused AdamW step -
    the whole thing captured in a CUDA graph and replayed (`use_graph`).

    `passes=2` reproduces the reference's two-pass video step (train.py:814-834, H3: loss = loss_0 + loss_1);
    throughput is reported per pass with passes=1.

    Gradient bookkeeping (reference: accelerator.accumulate / backward / optimizer.step / zero_grad, train.py:739-879):
      * gradients accumulate in the flat fp32 buffer across the micro-steps of a window; every loss is scaled by
        1 / accumulation (what accelerator.backward does);
      * the all-reduce runs only on the window's last micro-step;
      * with an attached optim.FusedAdamW the update kernel consumes and zeroes the gradients and rewrites the bf16 shadow,
        so neither a memset nor a cast pass remains in the step.  Without one (a torch optimizer, or fwd+bwd only) the
        buffer is zeroed at the start of each window and the shadow is re-cast from the masters every call."""

    def __init__(self, unet, alphas_cumprod, passes=1, use_graph=False, adopt=True, optimizer=None, accumulation=1):
        self.unet = unet
        self.abar = alphas_cumprod
        self.passes = passes
        self.arena = ParamArena(unet) if adopt else None
        self.use_graph = use_graph
        self.sync_gradients = True   # set False to run fwd+bwd only (profiling on a single rank)
        self.optimizer = optimizer
        self.accumulation = max(1, int(accumulation))
        self._micro = 0
        self._graphs = {}
        self._gscale = None
        # world > 1: per-block gradient all-reduces are issued from inside the backward pass (overlap), see GradientBuckets
        self.buckets = None
        if (self.arena is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
                and not os.environ.get("T2V_NO_OVERLAP")):
            self.buckets = GradientBuckets(self.arena, unet)
            self.buckets.install()

    def attach_optimizer(self, optimizer):
        self.optimizer = optimizer
        self._graphs = {}

    def _fwd_bwd(self, latents, noise, timesteps, text, first, last):
        """first / last: position of this micro-step inside its accumulation window."""
        fused = self.optimizer is not None and hasattr(self.optimizer, "launch")
        if self.arena is not None and not fused:
            if first:
                self.arena.zero_grads()
            self.arena.refresh_shadow()
        elif fused and first and not self.optimizer.covers_all_trainable():
            self.arena.zero_grads()   # trainable parameters outside the optimizer would otherwise accumulate forever
        ops.bump_dropout_epoch(latents.device)
        if self._gscale is None or self._gscale.device != latents.device:
            self._gscale = torch.empty((), device=latents.device, dtype=torch.float32)
            self._gscale.fill_(1.0 / self.accumulation)
        total = None
        reduce_now = last and self.sync_gradients
        overlap = self.buckets is not None and reduce_now
        for i in range(self.passes):
            loss = finetune_loss(self.unet, latents, noise, timesteps, text, self.abar)
            if overlap:
                self.buckets.armed = i == self.passes - 1   # gradients are final only in the last pass
            loss.backward(self._gscale if self.accumulation > 1 else None)
            total = loss.detach() if total is None else total + loss.detach()
        comm = None
        if overlap:
            self.buckets.finish()
            if fused:
                comm = self.buckets.comm          # bf16 all-reduced gradients: the update reads them directly
            else:
                self.buckets.widen()              # a torch optimizer / a test wants them in param.grad (fp32)
        elif reduce_now and self.arena is not None:
            allreduce_gradients(self.arena)
        if fused and last:
            self.optimizer.launch(zero_grad=True, grad_bf16=comm)
        return total

    def __call__(self, latents, noise, timesteps, encoder_hidden_states):
        if self.arena is not None:
            self.arena.check_layout()
            self.arena.reattach_grads()
        first = self._micro % self.accumulation == 0
        last = (self._micro + 1) % self.accumulation == 0
        self._micro += 1
        fused = self.optimizer is not None and hasattr(self.optimizer, "launch")
        if fused and last:
            self.optimizer.push_hyperparams()   # learning rate of this step -> device (outside the graph)
        args = (latents, noise, timesteps, encoder_hidden_states)
        if self.use_graph:
            key = (first, last, self.passes, self.sync_gradients, getattr(self.optimizer, "generation", 0),
                   tuple(tuple(a.shape) for a in args))
            g = self._graphs.get(key)
            if g is None:
                # the capture runs the step for real (warm-up + capture replays nothing): keep the optimizer state and the
                # weights of those dry runs out of the training trajectory
                g = self._graphs[key] = GraphedStep(lambda *a: self._fwd_bwd(*a, first, last), args, snapshot=self._snapshot())
            return g(*args)
        return self._fwd_bwd(*args, first, last)

    def _snapshot(self):
        """Tensors the dry runs of a graph capture must not change for good: weights, gradients and optimizer state."""
        keep = []
        if self.arena is not None:
            keep += [self.arena.master, self.arena.grad, self.arena.shadow]
        opt = self.optimizer
        if opt is not None and hasattr(opt, "launch"):
            keep += [opt.exp_avg, opt.exp_avg_sq, opt.state_dev, opt.sq]
        keep.append(ops.dropout_epoch(self.abar.device))
        return keep

    @property
    def _graph(self):   # bench.py / tests poke at this to drop captured graphs before tearing NCCL down
        return self._graphs

    @_graph.setter
    def _graph(self, value):
        self._graphs = {} if value is None else value
