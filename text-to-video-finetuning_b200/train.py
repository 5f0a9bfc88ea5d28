"""`python train.py --config cfg.yaml` - the reference's training entry point on the B200-native hot path.

`main(**yaml)` accepts the same keyword surface as the reference's `train.py:457-514`, so the v2 YAML configs load
unchanged (keys this build does not act on - validation sampling, webui export, trackers - are accepted and ignored).
What this file owns is the *step*: parameter selection (`handle_trainable_modules`, :316-337), LoRA injection through
`LoraHandler` (:557-572), optimizer parameter groups (`create_optimizer_params`, :205-236), and per optimisation step:
noise + timestep sampling (:751-757), the fused add_noise -> UNet fwd+bwd -> MSE step (`step.DataParallelStep`, two
passes per video step like :814-834), ONE gradient all-reduce across ranks, clipping, AdamW, LR schedule, LoRA / UNet
checkpoints.  One process per GPU (`torchrun --nproc-per-node N train.py --config ...`); no accelerate.

Data: `dataset_types` in ('json', 'single_video', 'image', 'folder') build the reference's dataset
classes (utils/dataset.py: OpenCV decode -> ONE resize + normalise kernel on the GPU -> batched AutoencoderKL.encode);
`cache_latents: True` writes / `cached_latent_dir` reads the reference's latent cache (`cached_{i}.pt`, train.py:266-314);
`dataset_types: ['synthetic']` needs no files.  Prompts go through the frozen CLIP text encoder (text_encoder.py) with a
per-prompt embedding cache; a batch that already carries `text_embeds` skips it.
Checkpoints (8(f) row 4): LoRA in the cloneofsimo list format, the UNet in diffusers layout, and - when the pretrained folder
is a full pipeline - the complete pipeline directory (`save_pipe`, train.py:395-449), plus a validation sample every
`validation_steps` (sampling.py: DPM-Solver++ preview with the trained UNet in eval mode, train.py:908-958).
"""
import argparse
import itertools
import math
import os
import time
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from .models.unet_3d_condition import UNet3DConditionModel
from .step import DataParallelStep, ddpm_alphas_cumprod, sample_noise
from .utils.lora_handler import LORA_VERSIONS, LoraHandler

already_printed_trainables = False


def handle_trainable_modules(model, trainable_modules=None, is_enabled=True, negation=None):
    """requires_grad by name-substring match; 'all' unlocks everything; LoRA params are never touched here."""
    global already_printed_trainables
    if trainable_modules is None:
        return
    count = 0
    if any(name == "all" for name in trainable_modules):
        model.requires_grad_(True)
        count = len(list(model.parameters()))
    else:
        model.requires_grad_(False)
        for name, param in model.named_parameters():
            if "lora" not in name and any(tm in name for tm in trainable_modules):
                param.requires_grad_(is_enabled)
                count += 1
    if count > 0 and not already_printed_trainables:
        already_printed_trainables = True
        print(f"{count} params have been processed.")


def param_optim(model, condition, extra_params=None, is_lora=False, negation=None):
    extra_params = extra_params if extra_params and len(extra_params.keys()) > 0 else None
    return {"model": model, "condition": condition, "extra_params": extra_params, "is_lora": is_lora, "negation": negation}


def _group(name=None, param=None, lr=None, extra=None, params=None):
    g = {"params": params} if params is not None else {"name": name, "params": param, "lr": lr}
    if extra:
        g.update(extra)
    return g


def create_optimizer_params(model_list, lr):
    groups = []
    for spec in model_list:
        model, condition, extra, is_lora, _ = spec.values()
        if not condition:
            continue
        if is_lora and isinstance(model, list):  # list of parameter iterators from the injector
            groups.append(_group(params=itertools.chain(*model), extra=extra))
        elif is_lora:
            groups += [_group(n, p, lr, extra) for n, p in model.named_parameters() if "lora" in n]
        else:
            groups += [_group(n, p, lr, extra) for n, p in model.named_parameters() if "lora" not in n]
    return groups


def _lr_lambda(kind, warmup, total):
    """diffusers.optimization.get_scheduler semantics (train.py:626-631): plain 'constant' has no warm-up."""
    def f(step):
        if kind == "constant":
            return 1.0
        if step < warmup:
            return float(step) / max(1, warmup)
        if kind == "constant_with_warmup":
            return 1.0
        prog = (step - warmup) / max(1, total - warmup)
        if kind == "linear":
            return max(0.0, 1.0 - prog)
        if kind == "cosine":
            return 0.5 * (1.0 + math.cos(math.pi * min(1.0, prog)))
        raise ValueError(f"unknown lr_scheduler {kind!r}")
    return f


class CachedLatents(torch.utils.data.Dataset):
    """The reference's latent cache: cached_{i}.pt = {pixel_values: latents (4,F,h,w), prompt_ids, text_prompt, ...}
    (train.py:266-314, utils/dataset.py:589-603) plus an optional precomputed 'text_embeds' (77, 1024) entry."""

    def __init__(self, cache_dir):
        self.files = sorted(os.path.join(cache_dir, f) for f in os.listdir(cache_dir) if f.endswith(".pt"))
        if not self.files:
            raise FileNotFoundError(f"no cached_*.pt latents under {cache_dir}")

    def __len__(self):
        return len(self.files)

    def __getitem__(self, i):
        return torch.load(self.files[i], map_location="cpu")


class SyntheticLatents(torch.utils.data.Dataset):
    def __init__(self, n=64, frames=16, hw=(32, 32), text_len=77, text_dim=1024, seed=0):
        self.n, self.shape, self.tshape, self.seed = n, (4, frames, hw[0], hw[1]), (text_len, text_dim), seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 100003 + i)
        return {"pixel_values": torch.randn(self.shape, generator=g) * 0.18215, "text_embeds": torch.randn(self.tshape, generator=g)}


def handle_cache_latents(should_cache, output_dir, train_dataloader, vae, device, cached_latent_dir=None):
    """reference train.py:266-314: encode every item once and store it as cached_{i}.pt = {pixel_values: latents (4, F, h, w)
    fp16, prompt_ids, text_prompt, dataset} (the reference's `batch[k] = v[0]` item format).  Returns the cache directory
    (None when caching is off); an existing `cached_latent_dir` is used as is."""
    if not should_cache:
        return None
    if cached_latent_dir is not None:
        return os.path.abspath(cached_latent_dir)
    from .utils.dataset import frames_to_latents
    cache_save_dir = os.path.join(output_dir, "cached_latents")
    os.makedirs(cache_save_dir, exist_ok=True)
    for i, batch in enumerate(train_dataloader):
        latents = frames_to_latents(batch, vae, device)
        item = {"pixel_values": latents[0].to(torch.float16).cpu()}
        for k, v in batch.items():
            if k in ("frames_u8", "pixel_values", "pixel_hw"):
                continue
            v0 = v[0]
            item[k] = v0.reshape(-1) if (torch.is_tensor(v0) and k == "prompt_ids") else v0
        torch.save(item, os.path.join(cache_save_dir, f"cached_{i}.pt"))
    return cache_save_dir


class TextEmbedder:
    """prompt ids -> encoder_hidden_states through the frozen text encoder (train.py:784-790), cached per distinct prompt:
    a finetune run repeats a handful of prompts thousands of times."""

    def __init__(self, text_encoder, device, max_entries=4096):
        self.enc, self.device, self.cache, self.max = text_encoder, device, {}, max_entries

    def __call__(self, prompt_ids):
        ids = prompt_ids.reshape(-1, prompt_ids.shape[-1]).to(torch.int64).cpu()
        out = []
        for row in ids:
            key = row.numpy().tobytes()
            e = self.cache.get(key)
            if e is None:
                e = self.enc(row[None].to(self.device))[0][0]
                if len(self.cache) < self.max:
                    self.cache[key] = e
            out.append(e)
        return torch.stack(out)


def _load_optional(cls, root, subfolder):
    return cls.from_pretrained(root, subfolder=subfolder) if os.path.isfile(os.path.join(root, subfolder, "config.json")) else None


def main(
    pretrained_model_path: str,
    output_dir: str,
    train_data: Dict = None,
    validation_data: Dict = None,
    extra_train_data: list = [],
    dataset_types: Tuple[str] = ("json",),
    shuffle: bool = True,
    validation_steps: int = 100,
    trainable_modules: Tuple[str] = None,
    trainable_text_modules: Tuple[str] = None,
    extra_unet_params=None,
    extra_text_encoder_params=None,
    train_batch_size: int = 1,
    max_train_steps: int = 500,
    learning_rate: float = 5e-5,
    scale_lr: bool = False,
    lr_scheduler: str = "constant",
    lr_warmup_steps: int = 0,
    adam_beta1: float = 0.9,
    adam_beta2: float = 0.999,
    adam_weight_decay: float = 1e-2,
    adam_epsilon: float = 1e-08,
    max_grad_norm: float = 1.0,
    gradient_accumulation_steps: int = 1,
    gradient_checkpointing: bool = False,
    text_encoder_gradient_checkpointing: bool = False,
    checkpointing_steps: int = 500,
    resume_from_checkpoint: Optional[str] = None,
    resume_step: Optional[int] = None,
    mixed_precision: Optional[str] = "fp16",
    use_8bit_adam: bool = False,
    enable_xformers_memory_efficient_attention: bool = True,
    enable_torch_2_attn: bool = False,
    seed: Optional[int] = None,
    train_text_encoder: bool = False,
    use_offset_noise: bool = False,
    rescale_schedule: bool = False,
    offset_noise_strength: float = 0.1,
    extend_dataset: bool = False,
    cache_latents: bool = False,
    cached_latent_dir=None,
    lora_version: str = LORA_VERSIONS[0],
    save_lora_for_webui: bool = False,
    only_lora_for_webui: bool = False,
    lora_bias: str = "none",
    use_unet_lora: bool = False,
    use_text_lora: bool = False,
    unet_lora_modules: Tuple[str] = ("ResnetBlock2D",),
    text_encoder_lora_modules: Tuple[str] = ("CLIPEncoderLayer",),
    save_pretrained_model: bool = True,
    lora_rank: int = 16,
    lora_path: str = "",
    lora_unet_dropout: float = 0.1,
    lora_text_dropout: float = 0.1,
    logger_type: str = "tensorboard",
    **kwargs,
):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(kwargs.get("device") or f"cuda:{local}")   # "device" is a test hook (CPU runs use emulated primitives)
    if dev.type == "cuda":
        torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if train_text_encoder or use_text_lora:
        raise NotImplementedError("text-encoder training is outside the B200 hot path of this build (SURVEY 8(f) row 2)")
    if use_8bit_adam:
        raise NotImplementedError("bitsandbytes 8-bit Adam is not available; fused AdamW is SURVEY 8(f) row 1")
    if seed is not None:
        # model construction and LoRA initialisation (lora_down ~ N(0, 1/r)) must be identical on every rank: the reference
        # gets that from accelerate/DDP broadcasting rank 0's parameters at wrap time (train.py:661).  The per-rank stream
        # (noise, timesteps, dropout) is seeded after the model is built.
        torch.manual_seed(seed)
    if rank == 0:
        os.makedirs(output_dir, exist_ok=True)

    unet = UNet3DConditionModel.from_pretrained(pretrained_model_path, subfolder="unet")
    unet.requires_grad_(False)
    if scale_lr:
        learning_rate = learning_rate * gradient_accumulation_steps * train_batch_size * world

    lora_manager = LoraHandler(version=lora_version, use_unet_lora=use_unet_lora, use_text_lora=False,
                               save_for_webui=save_lora_for_webui, only_for_webui=only_lora_for_webui,
                               unet_replace_modules=list(unet_lora_modules),
                               text_encoder_replace_modules=list(text_encoder_lora_modules), lora_bias=lora_bias)
    unet_lora_params, unet_negation = lora_manager.add_lora_to_model(use_unet_lora, unet, lora_manager.unet_replace_modules,
                                                                     lora_unet_dropout, lora_path, r=lora_rank)
    unet = unet.to(dev)
    unet.train()
    if kwargs.get("eval_train", False):  # train.py:779-781
        unet.eval()
    handle_trainable_modules(unet, trainable_modules, is_enabled=True, negation=unet_negation)
    unet._set_gradient_checkpointing(gradient_checkpointing)

    extra_unet_params = extra_unet_params or {}
    groups = create_optimizer_params([
        param_optim(unet, trainable_modules is not None, extra_params=extra_unet_params, negation=unet_negation),
        param_optim(unet_lora_params, use_unet_lora, is_lora=True, extra_params={**{"lr": learning_rate}, **extra_unet_params}),
    ], learning_rate)

    abar = ddpm_alphas_cumprod(device=dev)
    use_graph = bool(kwargs.get("use_cuda_graph", dev.type == "cuda"))   # replay the whole step as one CUDA graph (static shapes)
    stepper = DataParallelStep(unet, abar, passes=2, use_graph=use_graph, accumulation=gradient_accumulation_steps)
    # parameters now live in the flat arena.  Every rank must start from rank 0's weights (DDP does this at wrap time).
    if world > 1:
        dist.broadcast(stepper.arena.master, src=0)
        stepper.arena.refresh_shadow()
    if seed is not None:
        torch.manual_seed(seed + rank)
    fused_adamw = bool(kwargs.get("fused_adamw", True))   # optim.FusedAdamW on the flat arena; False: torch AdamW
    if fused_adamw:
        from .optim import FusedAdamW
        optimizer = FusedAdamW(stepper.arena, groups, lr=learning_rate, betas=(adam_beta1, adam_beta2), weight_decay=adam_weight_decay,
                               eps=adam_epsilon, max_grad_norm=max_grad_norm)
        stepper.attach_optimizer(optimizer)   # the update (clip + AdamW + shadow refresh + grad zeroing) is part of the step graph
    else:
        optimizer = torch.optim.AdamW(groups, lr=learning_rate, betas=(adam_beta1, adam_beta2), weight_decay=adam_weight_decay,
                                      eps=adam_epsilon)
    # stepped once per OPTIMIZER step, so warm-up and total are counted in optimizer steps (the reference scales both by the
    # accumulation factor, train.py:626-631, because accelerate steps its scheduler on every micro-step)
    sched = torch.optim.lr_scheduler.LambdaLR(optimizer, _lr_lambda(lr_scheduler, lr_warmup_steps, max_train_steps))

    kinds = [dataset_types] if isinstance(dataset_types, str) else list(dataset_types)
    # frozen side models, loaded only when the pretrained folder has them (a UNet-only folder trains from latents + embeddings)
    vae = text_encoder = tokenizer = None
    if dev.type == "cuda" or kwargs.get("load_side_models"):
        from .text_encoder import CLIPTextModel
        from .vae import AutoencoderKL
        vae = _load_optional(AutoencoderKL, pretrained_model_path, "vae")
        text_encoder = _load_optional(CLIPTextModel, pretrained_model_path, "text_encoder")
        if vae is not None:
            vae = vae.to(dev).eval()
        if text_encoder is not None:
            text_encoder = text_encoder.to(dev).eval()
        if os.path.isdir(os.path.join(pretrained_model_path, "tokenizer")):
            from transformers import CLIPTokenizer
            tokenizer = CLIPTokenizer.from_pretrained(pretrained_model_path, subfolder="tokenizer")
    embed_text = TextEmbedder(text_encoder, dev) if text_encoder is not None else None
    if cached_latent_dir:
        dataset = CachedLatents(cached_latent_dir)
    elif "synthetic" in kinds:
        td = train_data or {}
        dataset = SyntheticLatents(n=td.get("n", 64), frames=td.get("n_sample_frames", 16),
                                   hw=(td.get("height", 256) // 8, td.get("width", 256) // 8),
                                   text_dim=unet.config.cross_attention_dim)
    else:
        from .utils.dataset import extend_datasets, get_train_dataset
        if vae is None:
            raise FileNotFoundError(f"dataset_types={kinds} need the VAE of the pipeline: {pretrained_model_path}/vae is missing")
        if tokenizer is None or text_encoder is None:
            raise FileNotFoundError(f"dataset_types={kinds} need {pretrained_model_path}/tokenizer and /text_encoder for the prompts")
        parts = get_train_dataset(kinds, train_data, tokenizer)
        for extra in extra_train_data or []:
            parts += get_train_dataset(kinds, extra, tokenizer)
        extend_datasets(parts, ["train_data", "frames", "image_dir", "video_files"], extend=extend_dataset)
        parts = [d for d in parts if len(d) > 0]
        if not parts:
            raise FileNotFoundError(f"dataset_types={kinds}: no training items found under {train_data}")
        dataset = torch.utils.data.ConcatDataset(parts)
        if cache_latents:   # encode once, then train from the cache (reference handle_cache_latents)
            if rank == 0:
                handle_cache_latents(True, output_dir, torch.utils.data.DataLoader(dataset, batch_size=1, shuffle=False), vae, dev)
            if world > 1:
                dist.barrier()
            dataset = CachedLatents(os.path.join(output_dir, "cached_latents"))
    sampler = torch.utils.data.distributed.DistributedSampler(dataset, world, rank, shuffle=shuffle) if world > 1 else None
    loader = torch.utils.data.DataLoader(dataset, batch_size=train_batch_size, shuffle=shuffle and sampler is None, sampler=sampler)

    global_step, micro, epoch = 0, 0, 0
    t0 = time.time()
    step_times, t_iter = [], time.perf_counter()
    while global_step < max_train_steps:
        if sampler is not None:
            sampler.set_epoch(epoch)   # a new shuffle every epoch
        epoch += 1
        for batch in loader:
            if "frames_u8" in batch or (batch["pixel_values"].dim() == 5 and batch["pixel_values"].shape[2] == 3
                                        and batch["pixel_values"].shape[1] != 4):
                from .utils.dataset import frames_to_latents
                latents = frames_to_latents(batch, vae, dev)        # raw clip -> resize/normalise kernel -> batched VAE encode
            else:
                latents = batch["pixel_values"].to(dev, torch.float32)
            if "text_embeds" in batch:
                text = batch["text_embeds"].to(dev, torch.float32)
            elif embed_text is not None and "prompt_ids" in batch:
                text = embed_text(batch["prompt_ids"])                # frozen CLIP text encoder, cached per prompt
            else:
                raise FileNotFoundError("batch has neither 'text_embeds' nor a text encoder to turn 'prompt_ids' into them "
                                        f"({pretrained_model_path}/text_encoder is missing)")
            noise = sample_noise(latents, offset_noise_strength, use_offset_noise and not rescale_schedule)
            timesteps = torch.randint(0, abar.shape[0], (latents.shape[0],), device=dev, dtype=torch.int64)
            if latents.shape[2] <= 1:
                stepper.passes = 1  # single-frame data breaks out after the first pass (train.py:832)
            loss = stepper(latents, noise, timesteps, text)   # fused: on a window boundary this includes clip + AdamW
            micro += 1
            if micro % gradient_accumulation_steps:
                continue  # gradients keep accumulating in the flat buffer (loss scaled by 1 / accumulation)
            if fused_adamw:
                optimizer._opt_called = True   # the update ran inside the step (graph); keeps LambdaLR's order check quiet
            else:
                if max_grad_norm is not None:
                    torch.nn.utils.clip_grad_norm_([p for g in optimizer.param_groups for p in g["params"] if p.grad is not None], max_grad_norm)
                optimizer.step()
            sched.step()
            global_step += 1
            if kwargs.get("_time_steps"):   # test hook: synchronous wall time of the WHOLE iteration (data, noise, step, optimizer)
                torch.cuda.synchronize()
                now = time.perf_counter()
                step_times.append(now - t_iter)
                t_iter = now
            if rank == 0 and (global_step % 10 == 0 or global_step == 1):
                print(f"step {global_step}/{max_train_steps} loss {loss.item():.5f} ({(time.time() - t0) / global_step:.3f} s/step)")
            if rank == 0 and global_step % checkpointing_steps == 0:
                save_checkpoint(unet, lora_manager, output_dir, global_step, use_unet_lora, save_pretrained_model,
                                pretrained_model_path=pretrained_model_path)
            if rank == 0 and validation_data and validation_steps and global_step % validation_steps == 0 and vae is not None \
                    and text_encoder is not None and tokenizer is not None and getattr(vae, "decoder", None) is not None:
                from .sampling import validation_sample
                validation_sample(unet, vae, text_encoder, tokenizer, validation_data, os.path.join(output_dir, "samples"), global_step,
                                  batch.get("text_prompt", [""])[0] if isinstance(batch.get("text_prompt"), (list, tuple)) else "", dev)
            if global_step >= max_train_steps:
                break
    if world > 1:
        dist.barrier()
    if rank == 0:
        save_checkpoint(unet, lora_manager, output_dir, global_step, use_unet_lora, save_pretrained_model, final=True,
                        pretrained_model_path=pretrained_model_path)
    return {"steps": global_step, "step_times": step_times, "stepper": stepper, "optimizer": optimizer}


PIPELINE_PARTS = ("vae", "text_encoder", "tokenizer", "scheduler")


def save_pipe(pretrained_model_path, unet, path):
    """reference save_pipe (train.py:395-449): a complete TextToVideoSDPipeline directory - the trained UNet in diffusers
    layout plus the frozen parts and model_index.json of the pretrained pipeline, so `from_pretrained(path)` of a diffusers
    pipeline (or train.main again) can load it.  A UNet-only pretrained folder yields a UNet-only checkpoint."""
    import shutil
    os.makedirs(path, exist_ok=True)
    unet.save_pretrained(os.path.join(path, "unet"))
    copied = []
    for part in PIPELINE_PARTS:
        src = os.path.join(pretrained_model_path, part)
        if os.path.isdir(src):
            shutil.copytree(src, os.path.join(path, part), dirs_exist_ok=True)
            copied.append(part)
    index = os.path.join(pretrained_model_path, "model_index.json")
    if os.path.isfile(index):
        shutil.copyfile(index, os.path.join(path, "model_index.json"))
    elif copied:
        import json
        with open(os.path.join(path, "model_index.json"), "w") as f:
            json.dump({"_class_name": "TextToVideoSDPipeline", "unet": ["diffusers", "UNet3DConditionModel"],
                       "vae": ["diffusers", "AutoencoderKL"], "text_encoder": ["transformers", "CLIPTextModel"],
                       "tokenizer": ["transformers", "CLIPTokenizer"], "scheduler": ["diffusers", "DDIMScheduler"]}, f, indent=2)
    return copied


def save_checkpoint(unet, lora_manager, output_dir, step, use_unet_lora, save_pretrained_model, final=False, pretrained_model_path=None):
    """LoRA in the cloneofsimo list format (`lora/<step>_unet.pt`) and, with save_pretrained_model, the pipeline directory."""
    path = output_dir if final else os.path.join(output_dir, f"checkpoint-{step}")
    os.makedirs(path, exist_ok=True)
    if use_unet_lora:   # the reference saves the LoRA files and (save_pretrained_model) the pipeline: train.py:908-958
        from .utils.lora import save_lora_weight
        os.makedirs(os.path.join(path, "lora"), exist_ok=True)
        save_lora_weight(unet, os.path.join(path, "lora", f"{step}_unet.pt"), lora_manager.unet_replace_modules)
    if save_pretrained_model:
        if pretrained_model_path is not None:
            save_pipe(pretrained_model_path, unet, path)
        else:
            unet.save_pretrained(os.path.join(path, "unet"))


def load_config(path):
    import yaml
    with open(path) as f:
        return yaml.safe_load(f)


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--config", type=str, default="./configs/v2/train_config.yaml")
    main(**load_config(parser.parse_args().config))
