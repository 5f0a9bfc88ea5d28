"""Data pipeline: the reference's dataset classes on OpenCV decoding, item format, caption handling, the
latent cache round trip (`handle_cache_latents` format, reference train.py:266-314) and the device-side resize + normalise
kernel against torch's bilinear interpolation."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import emulated_prims


def _write_video(path, n=20, hw=(48, 64), fps=8):
    import cv2
    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (hw[1], hw[0]))
    rng = np.random.default_rng(0)
    base = rng.integers(0, 255, (hw[0], hw[1], 3), dtype=np.uint8)
    for i in range(n):
        w.write(np.roll(base, 3 * i, axis=1))
    w.release()


class _Tok:
    model_max_length = 77

    def __call__(self, prompt, **kw):
        ids = torch.zeros((1, 77), dtype=torch.int64)
        for i, ch in enumerate(prompt[:77]):
            ids[0, i] = ord(ch) % 1000
        return type("R", (), {"input_ids": ids})()


def test_folder_json_single_and_image_datasets(tmp_path):
    import cv2
    from t2v_b200.utils import dataset as D
    vids = tmp_path / "vids"
    vids.mkdir()
    for i in range(2):
        _write_video(str(vids / f"v{i}.mp4"))
    (vids / "v0.txt").write_text("a cat")
    tok = _Tok()
    ds = D.VideoFolderDataset(tokenizer=tok, width=32, height=24, n_sample_frames=4, fps=8, path=str(vids), fallback_prompt="fallback")
    assert len(ds) == 2 and ds.__getname__() == "folder"
    it = ds[0]
    assert it["frames_u8"].shape == (4, 48, 64, 3) and it["frames_u8"].dtype == torch.uint8
    assert it["text_prompt"] == "a cat" and ds[1]["text_prompt"] == "fallback" and it["prompt_ids"].shape == (1, 77)
    assert tuple(it["pixel_hw"].tolist()) == (24, 32)
    # reference item format on request
    ds_cpu = D.VideoFolderDataset(tokenizer=tok, width=32, height=24, n_sample_frames=4, fps=8, path=str(vids), device_preprocess=False)
    pv = ds_cpu[0]["pixel_values"]
    assert pv.shape == (4, 3, 24, 32) and pv.min() >= -1.0 and pv.max() <= 1.0
    # json
    jd = {"data": [{"video_path": str(vids / "v0.mp4"), "data": [{"frame_index": 2, "prompt": "p0"}, {"frame_index": 5, "prompt": "p1"}]}]}
    jp = tmp_path / "d.json"
    jp.write_text(json.dumps(jd))
    dj = D.VideoJsonDataset(tokenizer=tok, width=32, height=24, n_sample_frames=3, json_path=str(jp))
    assert len(dj) == 2 and dj[1]["text_prompt"] == "p1" and dj[1]["frames_u8"].shape[0] == 3 and dj.__getname__() == "json"
    # single video
    sv = D.SingleVideoDataset(tokenizer=tok, width=32, height=24, n_sample_frames=4, frame_step=2, single_video_path=str(vids / "v1.mp4"),
                              single_video_prompt="one video")
    assert len(sv) >= 2 and sv[0]["frames_u8"].shape[0] == 4 and sv[0]["text_prompt"] == "one video"
    # images
    imgs = tmp_path / "imgs"
    imgs.mkdir()
    cv2.imwrite(str(imgs / "a.png"), np.full((40, 50, 3), 128, np.uint8))
    (imgs / "a.txt").write_text("grey")
    di = D.ImageDataset(tokenizer=tok, width=32, height=24, image_dir=str(imgs), fallback_prompt="fb")
    assert len(di) == 1 and di[0]["frames_u8"].shape == (1, 40, 50, 3) and di[0]["text_prompt"] == "grey"
    # the YAML's `train_data:` section maps onto the classes
    built = D.get_train_dataset(["folder", "image"], dict(width=32, height=24, n_sample_frames=4, path=str(vids), image_dir=str(imgs)), tok)
    assert [d.__getname__() for d in built] == ["folder", "image"]


def _vae(device):
    from t2v_b200.vae import AutoencoderKL
    torch.manual_seed(3)
    return AutoencoderKL(block_out_channels=(32, 32, 64, 64), layers_per_block=1).to(device).eval()


@pytest.mark.parametrize("device", ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_frames_to_latents_and_cache_round_trip(tmp_path, device):
    from oracle import ops_ref
    from t2v_b200 import train
    from t2v_b200.utils import dataset as D
    vids = tmp_path / "vids"
    vids.mkdir()
    for i in range(3):
        _write_video(str(vids / f"v{i}.mp4"), hw=(64, 64))
    ds = D.VideoFolderDataset(tokenizer=_Tok(), width=32, height=32, n_sample_frames=4, fps=8, path=str(vids), fallback_prompt="x")
    loader = torch.utils.data.DataLoader(ds, batch_size=1, shuffle=False)
    import contextlib
    ctx = emulated_prims() if device == "cpu" else contextlib.nullcontext()
    old = ops_ref.BF
    if device == "cpu":
        ops_ref.BF = torch.float32
    try:
        with ctx:
            vae = _vae(device)
            batch = next(iter(loader))
            lat = D.frames_to_latents(batch, vae, torch.device(device), generator=None)
            assert lat.shape == (1, 4, 4, 4, 4) and torch.isfinite(lat).all()
            cache_dir = train.handle_cache_latents(True, str(tmp_path / "out"), loader, vae, torch.device(device))
    finally:
        ops_ref.BF = old
    files = sorted(os.listdir(cache_dir))
    assert files == ["cached_0.pt", "cached_1.pt", "cached_2.pt"]
    item = D.CachedDataset(cache_dir)[0]
    assert item["pixel_values"].shape == (4, 4, 4, 4) and item["prompt_ids"].shape == (77,) and item["text_prompt"] == "x"
    assert item["dataset"] == "folder"


@pytest.mark.gpu
def test_device_resize_normalise_matches_torch_bilinear():
    from oracle import ops_ref
    from t2v_b200 import prims
    g = torch.Generator().manual_seed(0)
    fr = torch.randint(0, 256, (5, 90, 120, 3), generator=g, dtype=torch.uint8).cuda()
    for hw in ((64, 64), (90, 120), (48, 200)):
        got = prims.frames_u8_to_nhwc8(fr, hw).float()
        want = ops_ref.frames_u8_to_nhwc8(fr, hw).float()
        assert got.shape == want.shape and (got[..., 3:] == 0).all()
        assert (got - want).abs().max().item() < 2e-2     # bf16 rounding of values in [-1, 1]


def test_sensible_buckets_match_the_reference():
    """Aspect-ratio bucketing vs vectors produced by the reference's own utils/bucketing.py (tests/golden/make_golden_buckets.py)."""
    import json
    from t2v_b200.utils.dataset import sensible_buckets
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "buckets.json")) as f:
        g = json.load(f)
    assert len(g["cases"]) >= 90
    for mw, mh, w, h, ow, oh in g["cases"]:
        got = sensible_buckets(mw, mh, w, h)
        assert (int(got[0]), int(got[1])) == (ow, oh), ((mw, mh, w, h), got, (ow, oh))
