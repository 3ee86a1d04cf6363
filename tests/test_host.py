"""CPU tests of the host-side logic: C-ABI surface, drop-in module contract, synthetic inputs,
batch sharding over a world_size-2 gloo group. No GPU compute is invoked here."""
import ctypes
import os
import re
import subprocess
import sys
from pathlib import Path

import pytest
import torch

import stlt_b200
from stlt_b200 import Stlt, StltModelConfig
from stlt_b200 import lib as L
from stlt_b200.sharding import shard_batch, shard_bounds
from stlt_b200.synthetic import make_batch, make_raw_boxes, random_state_dict

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "stlt_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stlt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    lib = L.load_library()
    names = declared_symbols()
    assert len(names) >= 17
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/stlt_b200.h but not exported"
    assert sorted(L.SIGNATURES) == names, "ctypes SIGNATURES out of sync with the header"


def test_no_gpu_means_loud_failure_not_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.load_library()
    dims = L.StltDims(768, 12, 4, 8, 4, 174, 256, 5, 1e-12, 1e-5)
    handle = ctypes.c_void_p()
    rc = lib.stlt_create(ctypes.byref(dims), ctypes.byref(handle))
    assert rc == L.STLT_ERR_CUDA
    assert b"CUDA" in lib.stlt_last_error(None) or b"device" in lib.stlt_last_error(None)
    model = Stlt(StltModelConfig(num_classes=174, unique_categories=4))
    model.train(False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(make_batch(2, "something"))


def test_create_rejects_unsupported_dims():
    lib = L.load_library()
    dims = L.StltDims(512, 8, 4, 8, 4, 174, 256, 5, 1e-12, 1e-5)
    handle = ctypes.c_void_p()
    assert lib.stlt_create(ctypes.byref(dims), ctypes.byref(handle)) == L.STLT_ERR_INVALID
    with pytest.raises(ValueError):
        Stlt(StltModelConfig(num_classes=10, unique_categories=4, hidden_size=512))


def test_state_dict_contract():
    """174 entries, reference names/shapes/dtypes."""
    m = Stlt(StltModelConfig(num_classes=174, unique_categories=4))
    sd = m.state_dict()
    assert len(sd) == 174
    assert sum(p.numel() for p in m.parameters()) == 93_080_238
    bfe = "backbone.frames_embeddings."
    assert sd[bfe + "position_ids"].dtype == torch.int64 and tuple(sd[bfe + "position_ids"].shape) == (1, 256)
    assert tuple(sd[bfe + "layout_embedding.category_box_embeddings.category_embeddings.weight"].shape) == (4, 768)
    assert tuple(sd[bfe + "layout_embedding.category_box_embeddings.box_embedding.weight"].shape) == (768, 4)
    assert tuple(sd[bfe + "layout_embedding.category_box_embeddings.score_embeddings.weight"].shape) == (768, 1)
    assert sum(k.startswith(bfe + "layout_embedding.encoder_layer.") for k in sd) == 12  # orphan layer
    for stack, n in ((bfe + "layout_embedding.transformer.layers.", 4), ("backbone.transformer.layers.", 8)):
        for i in range(n):
            p = f"{stack}{i}."
            assert tuple(sd[p + "self_attn.in_proj_weight"].shape) == (2304, 768)
            assert tuple(sd[p + "self_attn.out_proj.weight"].shape) == (768, 768)
            assert tuple(sd[p + "linear1.weight"].shape) == (3072, 768)
            assert tuple(sd[p + "linear2.weight"].shape) == (768, 3072)
    assert tuple(sd["prediction_head.fc2.weight"].shape) == (174, 768)
    assert tuple(sd[bfe + "frame_type_embedding.weight"].shape) == (5, 768)
    # clones at init, padding rows zero at init (nn.TransformerEncoder / nn.Embedding semantics)
    l0, l3 = (bfe + f"layout_embedding.transformer.layers.{i}.linear1.weight" for i in (0, 3))
    assert torch.equal(sd[l0], sd[l3]) and torch.equal(sd[l0], sd[bfe + "layout_embedding.encoder_layer.linear1.weight"])
    assert sd[bfe + "frame_type_embedding.weight"][0].abs().sum() == 0
    assert m.logit_names == ("stlt",)
    assert len(m.backbone.state_dict()) == 168
    m2 = Stlt(StltModelConfig(num_classes=174, unique_categories=4))
    m2.load_state_dict(random_state_dict(sd, seed=5), strict=True)
    assert m.train(False) is m and not m.training


def test_training_mode_has_no_cpu_path_either():
    m = Stlt(StltModelConfig(num_classes=174, unique_categories=4))
    assert m.training
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(make_batch(1, "something"))


def test_flat_training_layout_groups_parameters_like_the_reference_optimizer():
    """plan_flat_layout: add_weight_decay's two groups (train_inference_utils.py:37-54); the decayed tensors are ordered by
    the backward stage that finishes their gradient; orphan tensors are not part of the optimizer."""
    from stlt_b200.training import backward_stage_of, plan_buckets, plan_flat_layout
    m = Stlt(StltModelConfig(num_classes=174, unique_categories=4))
    layout, segments, total, stage_ends = plan_flat_layout(m.named_parameters())
    names = [n for n, _, _ in layout]
    assert len(names) == len(set(names)) == 161  # 173 parameters - 12 orphan encoder_layer.* tensors
    assert not any(".encoder_layer." in n for n in names)
    offs = [o for _, _, o in layout]
    assert offs == sorted(offs) and all(o % 4 == 0 for o in offs)
    params = dict(m.named_parameters())
    ns, nt = 4, 8
    assert len(stage_ends) == ns + nt + 3 and stage_ends == sorted(stage_ends) and stage_ends[-1] == segments["d"][1]
    for n, p, o in layout:
        key = next(k for k, (a, b) in segments.items() if a <= o < b)
        assert key.endswith("nd") == (p.dim() == 1 or n.endswith(".bias")), n
        assert key.startswith("sc") == ("score_embeddings" in n), n
        if key == "d":  # inside the slice of its backward stage
            k = backward_stage_of(n, ns, nt)
            assert (stage_ends[k - 1] if k else 0) <= o and o + p.numel() <= stage_ends[k], n
    assert segments["d"][0] == 0 and segments["sc_d"][1] == total
    assert total >= sum(params[n].numel() for n in names)
    # stage order = order in which stlt_backward finishes the gradients
    assert backward_stage_of("prediction_head.fc2.weight", ns, nt) == 0
    assert backward_stage_of("backbone.transformer.layers.7.linear1.weight", ns, nt) == 1
    assert backward_stage_of("backbone.transformer.layers.0.linear1.weight", ns, nt) == 8
    assert backward_stage_of("backbone.frames_embeddings.position_embeddings.weight", ns, nt) == 9
    assert backward_stage_of("backbone.frames_embeddings.layout_embedding.transformer.layers.3.linear2.weight", ns, nt) == 10
    assert backward_stage_of("backbone.frames_embeddings.layout_embedding.transformer.layers.0.linear2.weight", ns, nt) == 13
    assert backward_stage_of(
        "backbone.frames_embeddings.layout_embedding.category_box_embeddings.box_embedding.weight", ns, nt) == 14
    # every scheme tiles [0, total) with contiguous buckets in stage order; "two" cuts after the temporal phase
    for scheme in ("one", "two", "per_layer"):
        buckets = plan_buckets(scheme, stage_ends, total, nt + 2)
        assert buckets[0][1] == 0 and buckets[-1][2] == total and buckets[-1][0] == ns + nt + 2
        assert all(b0[2] == b1[1] and b0[0] < b1[0] for b0, b1 in zip(buckets, buckets[1:]))
        assert all(e > s_ for _, s_, e in buckets)
    assert len(plan_buckets("one", stage_ends, total, nt + 2)) == 1
    two = plan_buckets("two", stage_ends, total, nt + 2)
    assert len(two) == 2 and two[0] == (nt + 1, 0, stage_ends[nt + 1])
    per_layer = plan_buckets("per_layer", stage_ends, total, nt + 2)
    # one per encoder layer (the head and the frame embedding ride with the next layer) + the tail: category / box embedding,
    # the no-decay block and the score embedding (0.5 MB), the only bucket that waits for the very end of the backward pass
    assert len(per_layer) == 13 and per_layer[-1][2] - per_layer[-1][1] < 1 << 18
    with pytest.raises(ValueError):
        plan_buckets("three", stage_ends, total, nt + 2)


@pytest.mark.parametrize("layout,S,has_scores", [("something", 5, False), ("action_genome", 11, True)])
def test_synthetic_batch_follows_reference_layout_rules(layout, S, has_scores):
    spec = stlt_b200.SOMETHING_ELSE if layout == "something" else stlt_b200.ACTION_GENOME
    b = make_batch(16, layout, ragged=True, seed=1)
    B, L, S_ = b["categories"].shape
    assert (L, S_) == (17, S) and ("scores" in b) == has_scores
    assert (b["categories"][:, :, 0] == spec["cls_id"]).all()
    assert torch.equal(b["boxes"][:, :, 0], torch.tensor([0.0, 0.0, 1.0, 1.0]).expand(B, L, 4))
    for i in range(B):
        n = int(b["lengths"][i])
        assert b["frame_types"][i, n - 1] == spec["frame_types"]["extract"]
        assert (b["frame_types"][i, n:] == 0).all() and (b["frame_types"][i, :n] != 0).all()
        assert (b["categories"][i, n - 1:, 1:] == 0).all()
    assert (b["boxes"][b["categories"] == 0] == 0).all()
    assert torch.equal(b["src_key_padding_mask_boxes"], b["categories"] == 0)
    dense = make_batch(4, layout, ragged=False)
    assert (dense["lengths"] == 17).all() and (dense["categories"][:, :16] != 0).all()
    raw, sizes = make_raw_boxes(b["categories"], seed=2)
    assert raw.dtype == torch.float64 and tuple(sizes.shape) == (B, 2)


def test_shard_bounds_cover_batch():
    for B in (0, 1, 7, 8, 4096, 4099):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["STLT_ROOT"])
from stlt_b200.sharding import shard_batch, shard_bounds, gather_logits
from stlt_b200.synthetic import make_batch
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
batch = make_batch(7, "something", seed=4)
local = shard_batch(batch, world, rank)
lo, hi = shard_bounds(7, world, rank)
assert local["categories"].shape[0] == hi - lo
# stand-in for per-rank logits: a deterministic function of the local inputs
fake = local["boxes"].sum(dim=(1, 2)).float()          # [b_local, 4]
full = gather_logits(fake, 7)
want = batch["boxes"].sum(dim=(1, 2)).float()
assert torch.equal(full, want), (rank, full, want)
dist.barrier()
if rank == 0:
    print("SHARD_OK")
"""


def test_two_rank_gloo_sharding_roundtrip(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, STLT_ROOT=str(ROOT), MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "SHARD_OK" in res.stdout


_TRAIN_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["STLT_ROOT"])
import stlt_b200
from stlt_b200.training import all_reduce_buckets, plan_buckets, plan_flat_layout
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
model = stlt_b200.Stlt(stlt_b200.StltModelConfig(num_classes=174, unique_categories=4, num_spatial_layers=1,
                                                 num_temporal_layers=1))
layout, segments, total, stage_ends = plan_flat_layout(model.named_parameters())
# per-rank "gradients": the local-mean gradient already scaled by 1/world, as FusedTrainStep feeds the
# all-reduce (stlt_loss grad_scale = 1/world), so the SUM over ranks is the global-batch mean
want = sum(torch.randn(total, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
for scheme in ("one", "two", "per_layer"):
    local = torch.randn(total, generator=torch.Generator().manual_seed(100 + rank))
    flat = (local / world).clone()
    buckets = plan_buckets(scheme, stage_ends, total, 1 + 2)
    seen = []
    for w in all_reduce_buckets(flat, buckets, before_bucket=seen.append):
        w.wait()
    assert seen == [b[0] for b in buckets] and seen == sorted(seen), (scheme, seen)
    # every trainable, non-orphan tensor lies inside exactly one bucket and holds the global mean
    assert torch.allclose(flat, want, atol=1e-6), (rank, scheme)
for name, p, off in layout:
    assert 0 <= off and off + p.numel() <= total
dist.barrier()
if rank == 0:
    print("ALLREDUCE_OK")
"""


def test_two_rank_gloo_gradient_buckets(tmp_path):
    script = tmp_path / "train_worker.py"
    script.write_text(_TRAIN_WORKER)
    env = dict(os.environ, STLT_ROOT=str(ROOT), MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29519", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "ALLREDUCE_OK" in res.stdout


def test_embedding_layernorm_statistics_identity():
    """The identity embed_stats_kernel / embed_kernel rely on (csrc/elementwise.cu): the pre-LayerNorm row of
    CategoryBoxEmbeddings (models.py:29-39) is affine in u = (box, score, 1), so its mean is removed by centring the
    tables and its biased variance is u^T G_cat u / 768 with G_cat the Gram matrix of the centred vectors — checked
    here in numpy against the direct two-pass statistics (the CUDA kernels are checked on the GPU against the
    reference's embedding activations)."""
    import numpy as np
    rng = np.random.default_rng(0)
    H, U = 768, 7
    E = rng.normal(size=(U, H))
    W = rng.uniform(-0.5, 0.5, size=(H, 4))
    b = 0.02 * rng.normal(size=H)
    sw = rng.uniform(-1, 1, size=H)
    sb = 0.02 * rng.normal(size=H)
    for cat in range(U):
        for _ in range(5):
            box = rng.uniform(0, 1, size=4)
            score = rng.uniform(0.5, 1)
            x = E[cat] + W @ box + b + score * sw + sb
            mean, var = x.mean(), x.var()
            V = np.stack([W[:, 0], W[:, 1], W[:, 2], W[:, 3], sw, E[cat] + b + sb])  # [6, H]
            Vc = V - V.mean(axis=1, keepdims=True)
            G = Vc @ Vc.T / H
            u = np.array([*box, score, 1.0])
            assert abs(V.mean(axis=1) @ u - mean) < 1e-12
            assert abs(u @ G @ u - var) < 1e-12 * max(1.0, var)
            # the upper-triangle form the kernel evaluates (off-diagonal coefficients doubled)
            tri = sum((G[i, j] * (1.0 if i == j else 2.0)) * u[i] * u[j] for i in range(6) for j in range(i, 6))
            assert abs(tri - var) < 1e-12 * max(1.0, var)


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    """`bench.py --impl reference` launched like the B200 arm (torchrun, N = 2): rank 0 alone runs the CPU arm and
    prints ONE JSON line carrying the B200 arm's metric / unit / workload; the other rank exits 0 without work. The
    line says what ran: the unmodified reference from oracle/_ref when it has been built (kind "reference"), and the
    number of videos every timed step really forwarded."""
    import json
    from oracle import build_ref, ref_loader
    build_ref.build()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="4")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29523", str(ROOT / "bench.py"),
           "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--ref-batch", "16", "--no-extras"]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, res.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "stlt_inference_videos_per_sec" and d["unit"] == "videos/s"
    assert d["n_gpus"] == 2 and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"].startswith("STLT inference, something shape (L=17 frames x S=5 slots, 174 classes), batch 4096")
    assert d["config"]["executed_batch_per_step"] == 16 and "16 videos" in d["config"]["reference_sample"]
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port")
    assert d["cpu_baseline"]["value"] == d["value"]
    assert {(r["batch"], r["threads"]) for r in d["cpu_baseline"]["rows"]} >= {(8, 1), (16, d["cpu_baseline"]["cores"])}
    assert d["e2e"] == {"value": d["value"], "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_oracle_ref_is_the_unmodified_reference_and_agrees_with_the_port():
    """oracle/_ref (byte-compiled from /root/reference by oracle/build_ref.py) imports as the reference's own
    modelling.models and its Stlt forward equals the oracle restatement on re-drawn weights and a ragged batch."""
    import torch
    import stlt_b200
    from oracle import build_ref, ref_loader, stlt_oracle
    from stlt_b200.synthetic import make_batch, random_state_dict
    if not build_ref.build():
        import pytest
        pytest.skip("no /root/reference and no earlier oracle/_ref build on this machine")
    models, configs = ref_loader.load()
    assert "oracle/_ref/stlt_reference.bin" in models.__file__
    cfg = stlt_b200.StltModelConfig(num_classes=174, unique_categories=4)
    torch.manual_seed(0)
    sd = random_state_dict(stlt_b200.Stlt(cfg).state_dict(), seed=11)
    ref = models.Stlt(configs.StltModelConfig(num_classes=174, unique_categories=4))
    ref.load_state_dict(sd)  # strict: same 174 keys
    ref.train(False)
    batch = make_batch(6, "something", ragged=True, seed=12)
    with torch.no_grad():
        want = ref(batch)["stlt"]
        got = stlt_oracle.stlt_forward(sd, batch)
    assert float((got - want).abs().max() / want.abs().max()) < 2e-5


def test_verify_masks_accepts_collater_masks_and_rejects_custom_ones():
    import pytest
    import torch
    import stlt_b200
    from stlt_b200.synthetic import make_batch
    batch = make_batch(5, "something", ragged=True, seed=3)
    stlt_b200.Stlt.verify_masks(batch)  # the collater's masks: categories == 0, frame_types == 0
    bad = dict(batch)
    bad["src_key_padding_mask_boxes"] = batch["src_key_padding_mask_boxes"].clone()
    bad["src_key_padding_mask_boxes"][0, 0, 0] = True
    with pytest.raises(ValueError, match="custom padding masks"):
        stlt_b200.Stlt.verify_masks(bad)


def test_config_table_matches_reference_defaults():
    """The table-driven StltModelConfig carries the reference's field names and defaults (configs.py:92-111)."""
    import pytest
    import stlt_b200
    from oracle import build_ref, ref_loader
    cfg = stlt_b200.StltModelConfig(num_classes=174, unique_categories=4, some_future_kwarg=1)
    with pytest.raises(AssertionError):
        stlt_b200.StltModelConfig(unique_categories=4)
    if build_ref.build():
        _, configs = ref_loader.load()
        ref = configs.StltModelConfig(num_classes=174, unique_categories=4)
        for name in cfg.FIELDS:
            assert getattr(ref, name) == getattr(cfg, name), name


def test_bench_per_kernel_roofline_rows_and_frozen_layer_buckets():
    """Host logic without a GPU: (1) bench.roofline_by_kernel turns the per-role GEMM profile (stlt_get_profile_by_role)
    into fractions of the measured peaks — tensor peak for every role, HBM peak for the out-projection whose fused
    epilogue makes it bandwidth-bound; (2) the flat training layout keeps stlt_backward's stage numbering when whole
    layers are frozen (the layer counts come from the config, not from the trainable names)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("stlt_bench", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    peaks = {"bf16_tflops_sustained": 1000.0, "bf16_tflops": 1200.0, "hbm_gbs": 6000.0}
    rows_per_step = 1000 * 128
    roles = {"linear1": {"ms": 20.0, "flops": 2 * 2.0 * rows_per_step * 3072 * 768, "launches": 24},
             "out_proj": {"ms": 4.0, "flops": 2 * 2.0 * rows_per_step * 768 * 768, "launches": 24}}
    out = {r["role"]: r for r in bench.roofline_by_kernel(roles, 2, "bf16", peaks)}
    l1, op = out["linear1"], out["out_proj"]
    assert l1["launches_per_step"] == 12 and l1["ms_per_step"] == 10.0
    assert l1["achieved_tflops"] == pytest.approx(2.0 * rows_per_step * 3072 * 768 / 10e-3 / 1e12)
    assert l1["frac_of_tensor_peak"] == pytest.approx(l1["achieved_tflops"] / 1000.0) and "frac_of_hbm_peak" not in l1
    # out-projection, bf16: context in (2 B) + fp32 stream in and out (4 + 4) + bf16 operand copy out (2) per element
    assert op["algorithmic_gbps"] == pytest.approx(rows_per_step * 768 * 12 / 2e-3 / 1e9)
    assert op["frac_of_hbm_peak"] == pytest.approx(op["algorithmic_gbps"] / 6000.0)
    op32 = bench.roofline_by_kernel({"out_proj": roles["out_proj"]}, 2, "fp32", peaks)[0]
    assert op32["algorithmic_gbps"] == pytest.approx(rows_per_step * 768 * 16 / 2e-3 / 1e9)   # two planes each way
    assert op32["mma_frac_of_tensor_peak"] == pytest.approx(3 * op32["frac_of_tensor_peak"])

    from stlt_b200.training import plan_buckets, plan_flat_layout
    m = Stlt(StltModelConfig(num_classes=174, unique_categories=4))
    for n, p in m.named_parameters():
        p.requires_grad_(n.startswith("prediction_head.") or ".layers.7." in n)   # head + the last temporal layer
    layout, segments, total, stage_ends = plan_flat_layout(m.named_parameters(), 4, 8)
    assert len(stage_ends) == 15 and stage_ends[0] > 0 and stage_ends[1] > stage_ends[0]
    assert all(e == stage_ends[1] for e in stage_ends[1:])       # nothing trainable after stage 1
    two = plan_buckets("two", stage_ends, total, 10)
    assert two[0][0] == 9 and two[-1] == (14, stage_ends[9], total)   # the stage numbers the library's events use
    guessed = plan_flat_layout(m.named_parameters())[3]
    assert len(guessed) != 15   # read off the names alone the numbering would be wrong: hence the explicit counts
