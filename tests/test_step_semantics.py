"""Step-level semantics the advisor flagged in round 1, on CPU over the emulated primitives:
  * gradient accumulation: two micro-steps of one clip == one step on the two-clip batch (gradients and loss scaling);
  * dropout under gradient checkpointing: the recomputed forward draws the same masks, so checkpointed and plain runs give
    identical gradients;
  * the device epoch changes the masks from step to step (what keeps a replayed CUDA graph from repeating them);
  * two ranks running train.main with a seed and LoRA start from identical weights (2-rank gloo)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import emulated_prims, seeded_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINY = dict(block_out_channels=(32, 64, 64, 64), attention_head_dim=32, cross_attention_dim=32)


def _model(train=False):
    from t2v_b200.models.unet_3d_condition import UNet3DConditionModel
    m = UNet3DConditionModel(**TINY)
    m.load_state_dict(seeded_state_dict(m, 5))
    m.requires_grad_(True)
    return m.train() if train else m.eval()


def _inputs(B=2, F=2):
    g = torch.Generator().manual_seed(11)
    return (torch.randn(B, 4, F, 8, 8, generator=g), torch.randn(B, 4, F, 8, 8, generator=g), torch.tensor([100, 700][:B]),
            torch.randn(B, 3, 32, generator=g))


def test_accumulation_equals_batch():
    from oracle import leaves as L
    from oracle import ops_ref
    from t2v_b200 import step as S
    ops_ref.BF = torch.float32
    try:
        lat, noise, t, ehs = _inputs()
        with emulated_prims():
            m1 = _model()
            s1 = S.DataParallelStep(m1, L.ddpm_alphas_cumprod(), passes=1)
            loss_b = s1(lat, noise, t, ehs)
            g_batch = s1.arena.grad.clone()
            m2 = _model()
            s2 = S.DataParallelStep(m2, L.ddpm_alphas_cumprod(), passes=1, accumulation=2)
            la = s2(lat[:1], noise[:1], t[:1], ehs[:1])
            lb = s2(lat[1:], noise[1:], t[1:], ehs[1:])
            g_acc = s2.arena.grad.clone()
    finally:
        ops_ref.BF = torch.bfloat16
    assert abs(0.5 * (la + lb) - loss_b) < 1e-5 * abs(loss_b)
    assert (g_acc - g_batch).norm() < 1e-4 * g_batch.norm(), ((g_acc - g_batch).norm(), g_batch.norm())
    # a third call opens a new window: the buffer is zeroed first
    with emulated_prims():
        s2(lat[:1], noise[:1], t[:1], ehs[:1])
    assert (s2.arena.grad - g_batch).norm() > 1e-3 * g_batch.norm()


def test_dropout_masks_survive_checkpoint_recompute_and_change_per_step():
    from oracle import leaves as L
    from t2v_b200 import ops
    from t2v_b200 import step as S
    lat, noise, t, ehs = _inputs(B=1, F=2)
    grads = []
    with emulated_prims():
        for ckpt in (False, True):
            m = _model(train=True)            # TemporalConvLayer dropout p = 0.1 is live
            m._set_gradient_checkpointing(ckpt)
            torch.manual_seed(123)
            ops.dropout_epoch("cpu").zero_()   # both runs are "step 1"
            st = S.DataParallelStep(m, L.ddpm_alphas_cumprod(), passes=1)
            st(lat, noise, t, ehs)
            grads.append(st.arena.grad.clone())
        assert torch.equal(grads[0], grads[1]) or (grads[0] - grads[1]).norm() < 1e-6 * grads[0].norm()
        # same host seeds, next epoch -> different masks -> different gradients
        torch.manual_seed(123)
        st(lat, noise, t, ehs)
        assert (st.arena.grad - grads[1]).norm() > 1e-4 * grads[1].norm()


def _rank_main(rank, world, port, root, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0")
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import emulated_prims
    from t2v_b200 import train
    with emulated_prims():
        r = train.main(pretrained_model_path=root, output_dir=os.path.join(out, f"r{rank}"), dataset_types=["synthetic"],
                       train_data=dict(n=4, n_sample_frames=2, height=64, width=64), max_train_steps=1, learning_rate=1e-3,
                       checkpointing_steps=100, seed=64, device="cpu", use_unet_lora=True, lora_rank=4, lora_version="cloneofsimo",
                       unet_lora_modules=["UNet3DConditionModel"], trainable_modules=None, save_pretrained_model=False)
    torch.save(r["stepper"].arena.master.clone(), os.path.join(out, f"master{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_train_identical_lora_weights(tmp_path):
    from t2v_b200.models.unet_3d_condition import UNet3DConditionModel
    m = UNet3DConditionModel(**dict(TINY, block_out_channels=(64, 128, 128, 128), attention_head_dim=64, cross_attention_dim=64))
    m.load_state_dict(seeded_state_dict(m, 0))
    root = str(tmp_path / "model")
    m.save_pretrained(os.path.join(root, "unet"))
    out = str(tmp_path)
    mp.spawn(_rank_main, args=(2, 29600 + os.getpid() % 2000, root, out), nprocs=2, join=True)
    a, b = torch.load(os.path.join(out, "master0.pt")), torch.load(os.path.join(out, "master1.pt"))
    # identical LoRA initialisation on both ranks, identical (all-reduced) gradients, identical update
    assert torch.equal(a, b)


def test_fusions_are_wired_for_every_layer():
    """Host wiring of the two producer-side fusions, counted on a CPU run over the emulated primitives:
      * every GroupNorm forward of the UNet is handed the (sum, sum of squares) its producing GEMM accumulated - none computes
        its own statistics pass;
      * every biased conv / linear gets its bias gradient out of the weight-gradient launch (dbias argument); the column-sum
        primitive is left with the per-clip time-embedding gradients (rowbias layers) and the two channel-padded boundary convs."""
    from t2v_b200 import prims
    from t2v_b200 import step as S
    m = _model(train=False)
    lat, noise, t, text = _inputs(B=1, F=2)
    counts = {"gn": 0, "gn_with_stats": 0, "wgrad": 0, "wgrad_dbias": 0, "colsum": 0}
    with emulated_prims():
        gn0, wg0, cs0 = prims.groupnorm_fwd, prims.conv_wgrad, prims.colsum

        def gn(x, gamma, beta, G, eps, silu, stats=None, fps=1):
            counts["gn"] += 1
            counts["gn_with_stats"] += bool(stats)
            return gn0(x, gamma, beta, G, eps, silu, stats, fps)

        def wg(x, dy, dw, stride=1, pads=(0, 0, 0, 0), dbias=None):
            counts["wgrad"] += 1
            counts["wgrad_dbias"] += dbias is not None
            return wg0(x, dy, dw, stride, pads, dbias)

        def cs(x, out, Sn, P, C):
            counts["colsum"] += 1
            return cs0(x, out, Sn, P, C)

        prims.groupnorm_fwd, prims.conv_wgrad, prims.colsum = gn, wg, cs
        try:
            loss = S.finetune_loss(m, lat, noise, t, text, S.ddpm_alphas_cumprod())
            loss.backward()
        finally:
            prims.groupnorm_fwd, prims.conv_wgrad, prims.colsum = gn0, wg0, cs0
    biased = sum(1 for n, p in m.named_parameters() if n.endswith(".bias") and p.dim() == 1 and
                 n[:-5] + ".weight" in dict(m.named_parameters()) and dict(m.named_parameters())[n[:-5] + ".weight"].dim() >= 2)
    assert counts["gn"] > 50 and counts["gn_with_stats"] == counts["gn"], counts
    # rowbias layers: conv1 of every ResnetBlock2D (its bias gradient is the sum of the per-clip time-embedding gradient rows)
    n_resnets = sum(1 for mod in m.modules() if type(mod).__name__ == "ResnetBlock2D")
    # + conv_in / conv_out: their 4 latent channels are zero-padded to 8, the padded weight gradient goes through a temporary
    assert counts["colsum"] == n_resnets + 2, (counts, n_resnets)
    assert counts["wgrad_dbias"] >= biased - n_resnets - 4 and counts["wgrad_dbias"] > 0.5 * counts["wgrad"], (counts, biased, n_resnets)
    assert all(p.grad is not None for n, p in m.named_parameters() if n.endswith(".bias")), "a bias lost its gradient"
