"""Datasets of the finetune step - SURVEY 8(f) row 3 (reference utils/dataset.py, train.py:266-314).

Same class names, constructor keywords, `__getname__()` tags and item keys (`pixel_values`, `prompt_ids`, `text_prompt`,
`dataset`) as the reference's VideoJsonDataset / SingleVideoDataset / ImageDataset / VideoFolderDataset / CachedDataset, so
the `train_data:` section of the v2 YAML configs maps onto them unchanged.  What differs (B200-first):

  * decoding uses OpenCV (`cv2.VideoCapture`; decord is not available) and stops at RAW frames: an item carries
    `frames_u8` uint8 [F, H0, W0, 3] (RGB) and `pixel_hw`, the target size;
  * resize + normalisation run on the GPU in ONE kernel (`prims.frames_u8_to_nhwc8`: bilinear, x / 127.5 - 1, bf16
    channels-last - the layout AutoencoderKL.encode consumes), see `frames_to_latents`; all frames of a clip are encoded as one
    batch (the reference encodes frame by frame with vae slicing);
  * `pixel_values` (float [F, 3, h, w] in [-1, 1], the reference's item format) is still produced - on the CPU, with the same
    arithmetic - when a dataset is built with `device_preprocess=False` (tests, foreign consumers).
"""
import json
import os
import random
from glob import glob

import numpy as np
import torch
from torch.utils.data import Dataset

VID_TYPES = (".mp4", ".avi", ".mov", ".webm", ".flv", ".mjpeg")
IMG_TYPES = (".png", ".jpg", ".jpeg", ".bmp")


# ------------------------------------------------------------------------------------------------ helpers (reference :22-108)
def normalize_input(item, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), use_simple_norm=False):
    """uint8 / float frames [F, C, H, W] in 0..255 -> [-1, 1] (reference normalize_input: both of its branches reduce to
    x / 127.5 - 1 for the default mean = std = 0.5)."""
    item = item.float()
    if use_simple_norm or (tuple(mean) == (0.5, 0.5, 0.5) and tuple(std) == (0.5, 0.5, 0.5)):
        return item / 127.5 - 1.0
    m = torch.tensor(mean).view(1, -1, 1, 1)
    s = torch.tensor(std).view(1, -1, 1, 1)
    return (item / 255.0 - m) / s


def get_prompt_ids(prompt, tokenizer):
    return tokenizer(prompt, truncation=True, padding="max_length", max_length=tokenizer.model_max_length, return_tensors="pt").input_ids


def read_caption_file(caption_file):
    with open(caption_file, "r", encoding="utf8") as t:
        return t.read()


def get_text_prompt(text_prompt="", fallback_prompt="", file_path="", ext_types=(".mp4",), use_caption=False):
    """One caption file per media file (same stem, .txt) when use_caption is set; otherwise the given prompt."""
    try:
        if not use_caption:
            return text_prompt
        if len(text_prompt) > 1:
            return text_prompt
        for ext in ext_types:
            if file_path.endswith(ext):
                cand = file_path[:-len(ext)] + ".txt"
                if os.path.exists(cand):
                    return read_caption_file(cand)
        return fallback_prompt
    except OSError:
        print(f"Couldn't read prompt caption for {file_path}. Using fallback.")
        return fallback_prompt


def get_video_frames(n_total, start_idx, sample_rate=1, max_frames=24):
    frame_number = sorted((0, start_idx, n_total))[1]
    return list(range(frame_number, n_total, sample_rate))[:max_frames]


def sensible_buckets(m_width, m_height, w, h, min_size=192):
    """Aspect-ratio bucketing of the reference (utils/bucketing.py): keep one side, snap the other to a nearby bucket."""
    def closest(m_size, size):
        cands = [max(min_size, abs(int(m_size - m))) for m in (64, 128, 192)]
        return cands[min(range(len(cands)), key=lambda i: abs(cands[i] - size))]
    if h > w:
        return closest(m_width, m_width / (h / w)), m_height
    if h < w:
        return m_width, closest(m_height, m_height / (w / h))
    return m_width, m_height


class VideoReader:
    """Minimal decord.VideoReader stand-in on cv2.VideoCapture: len(), get_avg_fps(), get_batch(indices) -> uint8 [n, H, W, 3] RGB."""

    def __init__(self, path):
        import cv2
        self._cv2 = cv2
        self.path = path
        cap = cv2.VideoCapture(path)
        if not cap.isOpened():
            raise FileNotFoundError(f"cannot open video {path!r}")
        self._n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        self._fps = float(cap.get(cv2.CAP_PROP_FPS)) or 8.0
        cap.release()

    def __len__(self):
        return self._n

    def get_avg_fps(self):
        return self._fps

    def get_batch(self, indices):
        cv2 = self._cv2
        want = [int(i) for i in indices]
        cap = cv2.VideoCapture(self.path)
        frames, pos = {}, -1
        for idx in sorted(set(want)):
            if idx != pos + 1:
                cap.set(cv2.CAP_PROP_POS_FRAMES, idx)
            ok, f = cap.read()
            pos = idx
            if not ok:
                break
            frames[idx] = cv2.cvtColor(f, cv2.COLOR_BGR2RGB)
        cap.release()
        if not frames:
            raise RuntimeError(f"no frames decoded from {self.path!r}")
        last = frames[max(frames)]
        return torch.from_numpy(np.stack([frames.get(i, last) for i in want]))


def _read_image(path):
    import cv2
    img = cv2.imread(path, cv2.IMREAD_COLOR)
    if img is None:
        raise FileNotFoundError(path)
    return torch.from_numpy(cv2.cvtColor(img, cv2.COLOR_BGR2RGB))


def _cpu_pixel_values(frames_u8, hw):
    """The reference's item format: float [F, 3, h, w] in [-1, 1] (bilinear resize like the device kernel)."""
    x = frames_u8.permute(0, 3, 1, 2).float()
    if tuple(x.shape[-2:]) != tuple(hw):
        x = torch.nn.functional.interpolate(x, size=tuple(hw), mode="bilinear", align_corners=False)
    return normalize_input(x)


class _Base(Dataset):
    device_preprocess = True

    def _example(self, frames_u8, hw, prompt, prompt_ids):
        ex = {"prompt_ids": prompt_ids, "text_prompt": prompt, "dataset": self.__getname__(), "pixel_hw": torch.tensor(hw)}
        if self.device_preprocess:
            ex["frames_u8"] = frames_u8.contiguous()
        else:
            ex["pixel_values"] = _cpu_pixel_values(frames_u8, hw)
        return ex

    def _target_hw(self, frames_u8):
        if getattr(self, "use_bucketing", False):
            h0, w0 = frames_u8.shape[1:3]
            w, h = sensible_buckets(self.width, self.height, w0, h0)
            return int(h), int(w)
        return int(self.height), int(self.width)

    def _prompt_ids(self, prompt):
        return get_prompt_ids(prompt, self.tokenizer) if self.tokenizer is not None else torch.zeros((1, 77), dtype=torch.int64)


# ------------------------------------------------------------------------------------------------ datasets
class VideoJsonDataset(_Base):
    """JSON produced by the Video-BLIP2 preprocessor: {"data": [{"video_path": ..., "data": [{"frame_index", "prompt"}, ...]}]}
    (or per-clip entries with "clip_path" + "prompt")."""

    def __init__(self, tokenizer=None, width=256, height=256, n_sample_frames=4, sample_start_idx=1, frame_step=1, json_path="",
                 json_data=None, vid_data_key="video_path", preprocessed=False, use_bucketing=False, device_preprocess=True, **kwargs):
        self.tokenizer, self.use_bucketing, self.preprocessed, self.vid_data_key = tokenizer, use_bucketing, preprocessed, vid_data_key
        self.width, self.height, self.n_sample_frames = width, height, n_sample_frames
        self.sample_start_idx, self.frame_step, self.device_preprocess = sample_start_idx, frame_step, device_preprocess
        self.train_data = self.load_from_json(json_path, json_data)

    def load_from_json(self, path, json_data):
        try:
            if json_data is None:
                with open(path) as f:
                    json_data = json.load(f)
            out = []
            for data in json_data["data"]:
                for nested in data["data"]:
                    entry = {self.vid_data_key: data[self.vid_data_key], "frame_index": nested.get("frame_index", 0), "prompt": nested["prompt"]}
                    if nested.get("clip_path") is not None:
                        entry["clip_path"] = nested["clip_path"]
                    out.append(entry)
            return out
        except (OSError, KeyError, TypeError, ValueError):
            print("Non-existant JSON path. Skipping.")
            return None

    @staticmethod
    def __getname__():
        return "json"

    def __len__(self):
        return len(self.train_data) if self.train_data is not None else 0

    def __getitem__(self, index):
        d = self.train_data[index]
        path = d.get("clip_path") or d[self.vid_data_key]
        vr = VideoReader(path)
        start = 0 if d.get("clip_path") else d["frame_index"]
        idxs = get_video_frames(len(vr), start, self.frame_step, self.n_sample_frames)
        frames = vr.get_batch(idxs)
        prompt = d["prompt"]
        return self._example(frames, self._target_hw(frames), prompt, self._prompt_ids(prompt))


class SingleVideoDataset(_Base):
    """One video cut into consecutive chunks of n_sample_frames (every frame_step-th frame), one prompt for all."""

    def __init__(self, tokenizer=None, width=256, height=256, n_sample_frames=4, frame_step=1, single_video_path="", single_video_prompt="",
                 use_caption=False, use_bucketing=False, device_preprocess=True, **kwargs):
        self.tokenizer, self.use_bucketing, self.device_preprocess = tokenizer, use_bucketing, device_preprocess
        self.width, self.height, self.n_sample_frames, self.frame_step = width, height, n_sample_frames, frame_step
        self.single_video_path, self.single_video_prompt = single_video_path, single_video_prompt
        self.frames = []
        if os.path.exists(single_video_path):
            self.create_video_chunks()

    def create_video_chunks(self):
        n = len(VideoReader(self.single_video_path))
        idx = list(range(1, n, self.frame_step))
        chunks = [idx[i:i + self.n_sample_frames] for i in range(0, len(idx), self.n_sample_frames)]
        self.frames = [c for c in chunks if len(c) == self.n_sample_frames] or chunks[:1]
        return self.frames

    @staticmethod
    def __getname__():
        return "single_video"

    def __len__(self):
        return len(self.frames)

    def __getitem__(self, index):
        frames = VideoReader(self.single_video_path).get_batch(self.frames[index])
        prompt = self.single_video_prompt
        return self._example(frames, self._target_hw(frames), prompt, self._prompt_ids(prompt))


class ImageDataset(_Base):
    """A folder of images trained as single-frame clips; captions from `<image>.txt` or single_img_prompt."""

    def __init__(self, tokenizer=None, width=256, height=256, base_width=256, base_height=256, use_caption=False, image_dir="",
                 single_img_prompt="", use_bucketing=False, fallback_prompt="", device_preprocess=True, **kwargs):
        self.tokenizer, self.use_bucketing, self.device_preprocess = tokenizer, use_bucketing, device_preprocess
        self.width, self.height = width, height
        self.use_caption, self.single_img_prompt, self.fallback_prompt = use_caption, single_img_prompt, fallback_prompt
        self.image_dir = self.get_images_list(image_dir)

    def get_images_list(self, image_dir):
        if os.path.isdir(image_dir):
            return sorted(os.path.join(image_dir, x) for x in os.listdir(image_dir) if x.lower().endswith(IMG_TYPES))
        return []

    @staticmethod
    def __getname__():
        return "image"

    def __len__(self):
        return len(self.image_dir)

    def __getitem__(self, index):
        path = self.image_dir[index]
        frames = _read_image(path)[None]
        prompt = get_text_prompt(file_path=path, text_prompt=self.single_img_prompt, fallback_prompt=self.fallback_prompt,
                                 ext_types=IMG_TYPES, use_caption=True)
        return self._example(frames, self._target_hw(frames), prompt, self._prompt_ids(prompt))


class VideoFolderDataset(_Base):
    """A folder of .mp4 files with optional same-stem .txt captions; a random window of n_sample_frames at ~fps."""

    def __init__(self, tokenizer=None, width=256, height=256, n_sample_frames=16, fps=8, path="./data", fallback_prompt="",
                 use_bucketing=False, device_preprocess=True, **kwargs):
        self.tokenizer, self.use_bucketing, self.device_preprocess = tokenizer, use_bucketing, device_preprocess
        self.fallback_prompt = fallback_prompt
        self.video_files = sorted(glob(f"{path}/*.mp4"))
        self.width, self.height, self.n_sample_frames, self.fps = width, height, n_sample_frames, fps

    @staticmethod
    def __getname__():
        return "folder"

    def __len__(self):
        return len(self.video_files)

    def __getitem__(self, index):
        path = self.video_files[index]
        vr = VideoReader(path)
        n = self.n_sample_frames
        every = min(len(vr), max(1, round(vr.get_avg_fps() / self.fps)))
        eff = len(vr) // every
        n = min(n, eff)
        start = random.randint(0, eff - n)
        frames = vr.get_batch(every * np.arange(start, start + n))
        cap = path[:-4] + ".txt"
        prompt = read_caption_file(cap) if os.path.exists(cap) else self.fallback_prompt
        return self._example(frames, self._target_hw(frames), prompt, self._prompt_ids(prompt))


class CachedDataset(Dataset):
    """The latent cache written by handle_cache_latents (reference :589-603): cached_{i}.pt dicts with `pixel_values` = latents
    (4, F, h, w), `prompt_ids` (77,), `text_prompt`, `dataset`; an optional `text_embeds` (77, D) entry skips the text encoder."""

    def __init__(self, cache_dir=""):
        self.cache_dir = cache_dir
        self.cached_data_list = self.get_files_list()

    def get_files_list(self):
        return sorted(os.path.join(self.cache_dir, x) for x in os.listdir(self.cache_dir) if x.endswith(".pt"))

    def __len__(self):
        return len(self.cached_data_list)

    def __getitem__(self, index):
        return torch.load(self.cached_data_list[index], map_location="cpu", weights_only=False)


# ------------------------------------------------------------------------------------------------ device side
@torch.no_grad()
def frames_to_latents(batch, vae, device, generator=None):
    """A collated batch of raw clips -> latents (B, 4, F, h/8, w/8) * 0.18215 on `device`: H2D of the uint8 frames, one
    resize + normalise kernel, ONE batched VAE encode of all B*F frames, fused sample / rearrange / scale kernel
    (reference: normalize_input on the CPU, then tensor_to_vae_latent with per-frame slicing, train.py:339-347)."""
    from .. import prims
    if "frames_u8" in batch:
        fr = batch["frames_u8"]                       # [B, F, H0, W0, 3] uint8
        B, F = fr.shape[:2]
        hw = tuple(int(v) for v in batch["pixel_hw"].view(-1, 2)[0])
        x = fr.reshape((B * F,) + tuple(fr.shape[2:])).to(device, non_blocking=True).contiguous()
        nhwc8 = prims.frames_u8_to_nhwc8(x, hw)
    else:
        pv = batch["pixel_values"].to(device, torch.float32)   # [B, F, 3, h, w] in [-1, 1]
        B, F = pv.shape[:2]
        nhwc8 = prims.latents_to_nhwc8(pv.reshape(B * F, 3, 1, pv.shape[-2], pv.shape[-1]).contiguous())
    mom = vae.encode_moments_nhwc8(nhwc8)
    _, h, w, _ = mom.shape
    eps = torch.randn((B, 4, F, h, w), device=mom.device, dtype=torch.float32, generator=generator)
    return prims.vae_sample(mom, eps, B, F, 0.18215)


DATASETS = {cls.__getname__(): cls for cls in (VideoJsonDataset, SingleVideoDataset, ImageDataset, VideoFolderDataset)}


def get_train_dataset(dataset_types, train_data, tokenizer):
    """reference train.py `get_train_dataset`: one dataset per entry of dataset_types, built from the `train_data:` section."""
    out = []
    for kind in dataset_types:
        if kind not in DATASETS:
            raise ValueError(f"Dataset type not found: {kind} not in {sorted(DATASETS)}")
        out.append(DATASETS[kind](**dict(train_data or {}), tokenizer=tokenizer))
    if not out:
        raise ValueError("Dataset type not found: no dataset_types given")
    return out


def extend_datasets(datasets, dataset_items, extend=False):
    """reference train.py `extend_datasets`: repeat the shorter datasets' file lists up to the longest one."""
    biggest = max((len(d) for d in datasets), default=0)
    for d in datasets:
        for item in dataset_items:
            v = getattr(d, item, None)
            if v is None or not extend or len(v) == 0 or len(v) >= biggest:
                continue
            reps = biggest // len(v)
            setattr(d, item, (v * reps + v[:biggest - len(v) * reps]))
            print(f"New {d.__getname__()} dataset length: {len(d)}")
