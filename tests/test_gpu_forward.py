"""GPU parity tests of the whole path through the drop-in module / C ABI: CUDA vs the CPU oracle
and vs the golden fixtures produced by the reference. Tolerances are BASELINE.json's: fp32 logits
max|d|/max|ref| <= 1e-4, bf16 <= 2e-2 with identical top-1; masks and boxes bit-exact."""
import json

import numpy as np
import pytest
import torch

import stlt_b200
from oracle import stlt_oracle as O
from stlt_b200 import Stlt, StltModelConfig, prepare_layout_batch
from stlt_b200.synthetic import make_batch, make_raw_boxes, random_state_dict
from tests.util import golden_model_case, load_golden, nerr, to_cuda

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2


def _model(cfg, sd, precision):
    torch.manual_seed(0)
    m = Stlt(cfg, precision=precision)
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda")
    m.train(False)
    return m


# ---------------------------------------------------------------- K0: boxes + masks (bit exact)
def test_prepare_matches_reference_fix_box_golden():
    g = load_golden("fix_box.npz")
    n = g["raw"].shape[0]
    raw = torch.from_numpy(g["raw"]).view(n, 1, 1, 4).expand(n, 1, 2, 4).contiguous().cuda()
    cats = torch.tensor([[[3, 2]]]).expand(n, 1, 2).contiguous().cuda()
    out = prepare_layout_batch(raw, torch.from_numpy(g["sizes"]).cuda(), cats, torch.full((n, 1), 2).cuda())
    got = out["boxes"][:, 0, 1, :].cpu().numpy()
    assert np.array_equal(got.view(np.uint32), g["normalized"].view(np.uint32))
    assert torch.equal(out["boxes"][:, 0, 0, :].cpu(), torch.tensor([0.0, 0.0, 1.0, 1.0]).expand(n, 4))


@pytest.mark.parametrize("dataset", ["something", "action_genome"])
def test_prepare_matches_reference_collater_golden(dataset):
    """Raw JSON boxes scattered into the padded layout -> K0 -> must equal the reference
    StltDataset + StltCollater output bit for bit (boxes and both masks)."""
    g = load_golden(f"collate_{dataset}.npz")
    meta = json.loads(bytes(g["json"]).decode())
    cats = torch.from_numpy(g["categories"])
    B, L, S = cats.shape
    raw = torch.zeros(B, L, S, 4, dtype=torch.float64)
    sizes = torch.zeros(B, 2, dtype=torch.int64)
    for b, video in enumerate(meta["videos"]):
        sizes[b] = torch.tensor(meta["sizes"][video["id"]])
        idx = O.get_test_layout_indices(16, len(video["frames"]))
        for l, fi in enumerate(idx):
            objs = [e for e in video["frames"][fi]["frame_objects"] if e["score"] >= 0.5]
            for s, e in enumerate(objs, start=1):
                raw[b, l, s] = torch.tensor([e["x1"], e["y1"], e["x2"], e["y2"]], dtype=torch.float64)
    out = prepare_layout_batch(raw.cuda(), sizes.cuda(), cats.cuda(), torch.from_numpy(g["frame_types"]).cuda())
    assert np.array_equal(out["boxes"].cpu().numpy().view(np.uint32), g["boxes"].view(np.uint32))
    assert np.array_equal(out["src_key_padding_mask_boxes"].cpu().numpy(), g["src_key_padding_mask_boxes"])
    assert np.array_equal(out["src_key_padding_mask_frames"].cpu().numpy(), g["src_key_padding_mask_frames"])


def test_prepare_large_batch_matches_oracle_bit_exact():
    b = make_batch(4096, "something", ragged=True, seed=12)
    raw, sizes = make_raw_boxes(b["categories"], seed=13)
    want = O.prepare_padded(raw, sizes, b["categories"], b["frame_types"])
    got = prepare_layout_batch(raw.cuda(), sizes.cuda(), b["categories"].cuda(), b["frame_types"].cuda())
    for k in want:
        w, g_ = want[k].numpy(), got[k].cpu().numpy()
        assert np.array_equal(w.view(np.uint32) if w.dtype == np.float32 else w,
                              g_.view(np.uint32) if g_.dtype == np.float32 else g_), k
    empty = prepare_layout_batch(raw[:0].cuda(), sizes[:0].cuda(), b["categories"][:0].cuda(), b["frame_types"][:0].cuda())
    assert empty["boxes"].shape == (0, 17, 5, 4)


# ---------------------------------------------------------------- forward vs golden / oracle
@pytest.mark.parametrize("layout", ["something", "action_genome"])
def test_forward_fp32_matches_reference_golden_per_stage(layout):
    cfg, sd, batch, g = golden_model_case(layout)
    m = _model(cfg, sd, "fp32")
    with torch.no_grad():
        out = m.forward_with_taps(to_cuda(batch))
    assert torch.equal(out["src_key_padding_mask_boxes"].cpu(), batch["src_key_padding_mask_boxes"])
    assert torch.equal(out["src_key_padding_mask_frames"].cpu(), batch["src_key_padding_mask_frames"])
    assert nerr(out["embed"][0], torch.from_numpy(g["embed_b0"])) < 1e-5
    valid = ~batch["src_key_padding_mask_boxes"][0]
    assert nerr(out["spatial"][0].cpu()[valid], torch.from_numpy(g["spatial_b0"])[valid]) < FP32_TOL
    assert nerr(out["frames"], torch.from_numpy(g["frames"])) < FP32_TOL
    for b in range(batch["lengths"].shape[0]):
        n = int(batch["lengths"][b])
        assert nerr(out["temporal"][b, :n], torch.from_numpy(g["temporal"][b, :n])) < FP32_TOL
    err = nerr(out["stlt"], torch.from_numpy(g["logits"]))
    print(f"{layout} fp32 logits nerr {err:.3e}")
    assert err < FP32_TOL
    m.check_inputs()
    assert m.last_launch_count() > 80


@pytest.mark.parametrize("layout", ["something", "action_genome"])
def test_forward_bf16_matches_reference_golden(layout):
    cfg, sd, batch, g = golden_model_case(layout)
    m = _model(cfg, sd, "bf16")
    with torch.no_grad():
        got = m(to_cuda(batch))["stlt"].cpu()
    want = torch.from_numpy(g["logits"])
    err = nerr(got, want)
    print(f"{layout} bf16 logits nerr {err:.3e}")
    assert err < BF16_TOL


def test_forward_batch8_config1_fp32_and_bf16_top1():
    """BASELINE config 1 shape (batch 8) on independently drawn weights, both precisions."""
    cfg = StltModelConfig(num_classes=174, unique_categories=4)
    torch.manual_seed(0)
    sd = random_state_dict(Stlt(cfg).state_dict(), seed=31)
    batch = make_batch(8, "something", ragged=True, seed=32)
    with torch.no_grad():
        want = O.stlt_forward(sd, batch)
        want64 = O.stlt_forward(sd, batch, dtype=torch.float64)
    m = _model(cfg, sd, "fp32")
    with torch.no_grad():
        got = m(to_cuda(batch))["stlt"].cpu()
    assert nerr(got, want) < FP32_TOL and nerr(got, want64) < FP32_TOL
    assert torch.equal(got.argmax(-1), want.argmax(-1))
    m.precision = "bf16"
    with torch.no_grad():
        got16 = m(to_cuda(batch))["stlt"].cpu()
    assert nerr(got16, want) < BF16_TOL
    margin = want.topk(2, -1).values
    safe = (margin[:, 0] - margin[:, 1]) > 4 * (got16 - want).abs().max()
    assert torch.equal(got16.argmax(-1)[safe], want.argmax(-1)[safe])
    assert torch.equal(got16.argmax(-1), want.argmax(-1)), "bf16 top-1 differs on the fixed batch"
    m.precision = "fp32"  # switching back re-packs the weights
    with torch.no_grad():
        again = m(to_cuda(batch))["stlt"].cpu()
    assert torch.equal(again, got)


def test_eval_mode_uses_inference_path_with_or_without_no_grad():
    """Eval mode returns the configured-precision inference logits (no grad_fn) whether or not the caller wraps
    the call in torch.no_grad() (src/inference.py:60-77 does, an ad-hoc notebook call may not); train mode
    returns logits with a grad_fn (src/train.py:125-127)."""
    cfg = StltModelConfig(num_classes=174, unique_categories=4)
    torch.manual_seed(0)
    sd = random_state_dict(Stlt(cfg).state_dict(), seed=41)
    batch = to_cuda(make_batch(6, "something", ragged=True, seed=42))
    m = _model(cfg, sd, "fp32")
    with torch.no_grad():
        a = m(batch)["stlt"]
    b = m(batch)["stlt"]
    assert b.grad_fn is None and not b.requires_grad
    assert torch.equal(a, b)
    m.train(True)
    m.config.hidden_dropout_prob = 0.0
    c = m(batch)["stlt"]
    assert c.grad_fn is not None
    assert nerr(c.detach(), a) < BF16_TOL


def test_forward_default_init_clone_layers():
    """Default init (all encoder layers identical clones, zero padding rows), non-128-multiple M."""
    cfg = StltModelConfig(num_classes=174, unique_categories=4)
    torch.manual_seed(3)
    m = Stlt(cfg).to("cuda")
    m.train(False)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    batch = make_batch(5, "something", ragged=True, seed=33)
    with torch.no_grad():
        want = O.stlt_forward(sd, batch)
        got = m(to_cuda(batch))["stlt"].cpu()
    assert nerr(got, want) < FP32_TOL


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_last_layer_pruning_is_bit_identical(precision):
    """Only slot 0 / frame lengths-1 of the stack outputs are read; running the
    row-wise tail of the last layers on those rows only must not change a single bit."""
    cfg = StltModelConfig(num_classes=157, unique_categories=38)
    torch.manual_seed(0)
    sd = random_state_dict(Stlt(cfg).state_dict(), seed=61)
    m = _model(cfg, sd, precision)
    batch = to_cuda(make_batch(37, "action_genome", ragged=True, seed=62))
    m.set_fused_layer_norm(False)  # the LayerNorm-fused bf16 path only exists in its pruned form
    m.set_compaction(False)        # the pad-skipping layout moves rows to other tile positions: round-off, not bits (below)
    with torch.no_grad():
        pruned = m(batch)["stlt"].clone()
        n_pruned = m.last_launch_count()
        m.set_pruning(False)
        full = m(batch)["stlt"].clone()
        n_full = m.last_launch_count()
        m.set_pruning(True)
        m.set_compaction(True)
        compact = m(batch)["stlt"].clone()
        n_compact = m.last_launch_count()
    assert torch.equal(pruned, full)
    assert n_pruned == n_full + 1  # two gathers replace the final gather_last
    # the same separate kernels on the pad-skipping layout: 3 planning kernels + one more attention launch per spatial layer
    print(precision, "pad-skipping vs padded grid", nerr(compact, pruned), "launches", n_compact, n_pruned)
    assert n_compact == n_pruned + 3 + cfg.num_spatial_layers
    # bf16: operand rounding flips with the tile position of a row, so the two layouts agree at the bf16-noise level only
    assert nerr(compact, pruned) < (2e-5 if precision == "fp32" else 1.5e-2)


def test_weight_update_is_picked_up():
    cfg = StltModelConfig(num_classes=174, unique_categories=4)
    torch.manual_seed(0)
    sd = random_state_dict(Stlt(cfg).state_dict(), seed=41)
    m = _model(cfg, sd, "fp32")
    batch = to_cuda(make_batch(2, "something", seed=1))
    with torch.no_grad():
        a = m(batch)["stlt"].clone()
        m.backbone.transformer.layers[7].linear2.weight.mul_(0.5)     # in-place: version bump
        m.prediction_head.fc2.bias.add_(1.0)
        b = m(batch)["stlt"]
    sd2 = {k: v.cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        want = O.stlt_forward(sd2, {k: v.cpu() for k, v in batch.items()})
    assert not torch.allclose(a, b) and nerr(b, want) < FP32_TOL


def test_edge_cases_empty_single_and_bad_inputs():
    cfg = StltModelConfig(num_classes=174, unique_categories=4)
    torch.manual_seed(0)
    sd = random_state_dict(Stlt(cfg).state_dict(), seed=42)
    m = _model(cfg, sd, "fp32")
    empty = to_cuda(make_batch(0, "something"))
    with torch.no_grad():
        assert m(empty)["stlt"].shape == (0, 174)
    # a single video with the minimum length (1 real frame + extract), L = 2, S = 2
    one = make_batch(1, "something", ragged=False, seed=2, num_frames=1, max_objects=1)
    with torch.no_grad():
        want = O.stlt_forward(sd, one)
        got = m(to_cuda(one))["stlt"].cpu()
    assert nerr(got, want) < FP32_TOL
    # extra keys are tolerated (move_batch_to_device leaves non-tensors in place)
    b = to_cuda(make_batch(2, "something", seed=3))
    b["video_id"] = ["a", "b"]
    b["labels"] = torch.zeros(2, dtype=torch.int64, device="cuda")
    with torch.no_grad():
        m(b)
    # out-of-range category is reported (PyTorch raises IndexError on CPU)
    bad = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in b.items()}
    bad["categories"][0, 0, 1] = 99
    with torch.no_grad():
        m(bad)
    with pytest.raises(stlt_b200.lib.StltError, match="categories"):
        m.check_inputs()
    with pytest.raises(TypeError):
        m({**b, "categories": b["categories"].int()})
    with pytest.raises(ValueError):
        m({**b, "boxes": b["boxes"][:, :, :, :3]})


# ---------------------------------------------------------------- full-size properties (B = 4096)
def test_full_size_batch_properties():
    """BASELINE config 2 size. The oracle cannot run 4096 videos in seconds, so check
    size-independent properties: (1) a 64-video sub-batch run alone gives the same logits as inside
    the big batch; (2) scrambling padded slots / padded frames changes nothing;
    (3) a sample of videos matches the oracle within the fp32 tolerance."""
    cfg = StltModelConfig(num_classes=174, unique_categories=4)
    torch.manual_seed(0)
    sd = random_state_dict(Stlt(cfg).state_dict(), seed=51)
    m = _model(cfg, sd, "fp32")
    batch = make_batch(4096, "something", ragged=True, seed=52)
    gb = to_cuda(batch)
    with torch.no_grad():
        compact_full = m(gb)["stlt"].clone()   # default: pad-skipping layout (rows of a video depend on the videos before it)
        m.set_compaction(False)                 # the bit-exactness properties below are properties of the padded grid
        full = m(gb)["stlt"].clone()
        assert nerr(compact_full, full) < 2e-5
        assert torch.isfinite(full).all()
        sub = {k: v[1000:1064].contiguous() for k, v in gb.items()}
        part = m(sub)["stlt"]
        # Sequences share 32-row attention tiles, so a sequence's position inside its tile (and with it
        # the order of the tensor-core partial sums) depends on where the batch starts: equal to fp32
        # round-off, like the reference itself (~1e-6 across batch compositions).
        assert nerr(part, full[1000:1064]) < 2e-5, "logits depend on batch composition"
        aligned = {k: v[960:1152].contiguous() for k, v in gb.items()}  # 960*17 frames: tile aligned
        assert torch.equal(m(aligned)["stlt"], full[960:1152]), "aligned sub-batch is not bit-identical"
        scr = {k: v.clone() for k, v in gb.items()}
        scr["boxes"][scr["categories"] == 0] = 0.77
        pad_frames = scr["frame_types"] == 0
        scr["boxes"][pad_frames] = 0.123
        assert torch.equal(m(scr)["stlt"], full), "padded payload leaked into the logits"
    idx = torch.tensor([0, 1, 777, 2048, 4095])
    small = {k: v[idx] for k, v in batch.items()}
    with torch.no_grad():
        want = O.stlt_forward(sd, small)
    assert nerr(full[idx.cuda()].cpu(), want) < FP32_TOL
    assert nerr(compact_full[idx.cuda()].cpu(), want) < FP32_TOL
    with torch.no_grad():  # padded payload cannot leak on the pad-skipping layout either (bit-exact: same rows, same tiles)
        m.set_compaction(True)
        assert torch.equal(m(scr)["stlt"], compact_full)
    m.precision = "bf16"
    with torch.no_grad():
        b16 = m(gb)["stlt"]
    assert nerr(b16[idx.cuda()].cpu(), want) < BF16_TOL


def test_host_pipeline_delivers_every_batch_in_order():
    """stlt_b200.pipeline.HostPipeline overlaps H2D / compute / D2H of neighbouring batches; results must
    be the ones of the plain sequential loop, in order, for ragged batch sizes too."""
    import stlt_b200
    from stlt_b200.pipeline import HostPipeline
    from stlt_b200.synthetic import make_batch, random_state_dict
    cfg = stlt_b200.StltModelConfig(num_classes=174, unique_categories=4, num_spatial_layers=1, num_temporal_layers=1)
    torch.manual_seed(0)
    model = stlt_b200.Stlt(cfg, precision="bf16")
    model.load_state_dict(random_state_dict(model.state_dict(), seed=3))
    model = model.cuda()
    model.train(False)
    sizes = [5, 9, 9, 2, 16, 16, 16, 1]
    keys = ("categories", "boxes", "frame_types", "lengths")
    host = [{k: v.pin_memory() for k, v in make_batch(b, "something", seed=40 + i).items() if k in keys}
            for i, b in enumerate(sizes)]
    with torch.no_grad():
        want = [model({k: v.cuda() for k, v in hb.items()})["stlt"].cpu() for hb in host]
    pipe = HostPipeline(model, "stlt", depth=2)
    got = [t.clone() for t in pipe.run(iter(host))]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    got3 = [t.clone() for t in HostPipeline(model, "stlt", depth=3).run(iter(host))]
    assert all(torch.equal(a, b) for a, b in zip(got3, want))


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_cuda_graph_replay_matches_eager(precision):
    """The forward is CUDA-graph capturable (no allocation / sync inside the library): a replayed graph
    gives bit-identical logits for new inputs of the captured shape."""
    import stlt_b200
    from stlt_b200.synthetic import make_batch, random_state_dict
    cfg = stlt_b200.StltModelConfig(num_classes=174, unique_categories=4)
    torch.manual_seed(0)
    model = stlt_b200.Stlt(cfg, precision=precision)
    model.load_state_dict(random_state_dict(model.state_dict(), seed=5))
    model = model.cuda()
    model.train(False)
    keys = ("categories", "boxes", "frame_types", "lengths")
    b0 = {k: v.cuda() for k, v in make_batch(8, "something", ragged=False, seed=1).items() if k in keys}
    run = model.make_graphed(b0)
    for seed in (2, 3):
        b = {k: v.cuda() for k, v in make_batch(8, "something", ragged=True, seed=seed).items() if k in keys}
        b["lengths"][0] = 17
        with torch.no_grad():
            want = model(b)["stlt"].clone()
        got = run(b)["stlt"]
        torch.cuda.synchronize()
        assert torch.equal(got, want)
    # latency of the graphed batch-8 forward (informational)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(50):
        run(b0)
    end.record()
    torch.cuda.synchronize()
    g_ms = start.elapsed_time(end) / 50
    start.record()
    with torch.no_grad():
        for _ in range(50):
            model(b0)
    end.record()
    torch.cuda.synchronize()
    print(f"batch-8 forward {precision}: graph {g_ms:.3f} ms, eager {start.elapsed_time(end) / 50:.3f} ms")


@pytest.mark.parametrize("layout", ["something", "action_genome"])
def test_fused_layer_norm_path_matches_unfused_and_oracle(layout):
    """bf16 mode folds the encoder LayerNorms into the GEMM epilogues (pre-norm residual stream + row
    statistics). A/B against the separate add+LN kernels and against the oracle on the golden case."""
    import stlt_b200
    from oracle import stlt_oracle
    from tests.util import golden_model_case, nerr, to_cuda
    cfg, sd, batch, g = golden_model_case(layout)
    model = stlt_b200.Stlt(cfg, precision="bf16")
    model.load_state_dict(sd)
    model = model.cuda()
    model.train(False)
    gb = to_cuda(batch)
    with torch.no_grad():
        model.set_fused_layer_norm(True)
        comp = model(gb)["stlt"].float().cpu()   # default: LayerNorms + attention in the GEMM epilogues, pad rows skipped
        n_comp = model.last_launch_count()
        model.set_hilo_residual(True)
        hilo = model(gb)["stlt"].float().cpu()   # optional: residual stream as two bf16 planes instead of fp32 + bf16 copy
        n_hilo = model.last_launch_count()
        model.set_hilo_residual(False)
        model.set_compaction(False)
        full = model(gb)["stlt"].float().cpu()   # the same on the whole padded [B, L, S] grid
        n_full = model.last_launch_count()
        model.set_fused_attention(False)
        fused = model(gb)["stlt"].float().cpu()
        n_fused = model.last_launch_count()
        model.set_fused_layer_norm(False)
        plain = model(gb)["stlt"].float().cpu()
        n_plain = model.last_launch_count()
        want = stlt_oracle.stlt_forward(sd, batch)
    print(layout, "fused vs oracle", nerr(fused, want), "plain vs oracle", nerr(plain, want), "fused vs plain",
          nerr(fused, plain), "launches", n_fused, n_plain)
    assert n_fused == n_plain - 23  # 24 residual + LayerNorm launches gone, one LayerNorm of the pooled rows added
    assert n_full == n_fused - 12   # the 12 attention launches live in the in-projection epilogues
    assert n_comp == n_full + 3     # three tiny planning kernels of the pad-skipping layout
    # bf16 rounding of the GEMM operands flips on 2^-18 perturbations of the stream: equal at the bf16-noise level only
    assert n_hilo == n_comp and nerr(hilo, want) < 2e-2 and nerr(comp, hilo) < 1.5e-2
    print(layout, "pad-skipping vs oracle", nerr(comp, want), "vs padded grid", nerr(comp, full))
    assert nerr(comp, want) < 2e-2 and nerr(comp, full) < 5e-3 and torch.equal(comp.argmax(-1), want.argmax(-1))
    print(layout, "attention-fused vs oracle", nerr(full, want), "vs LayerNorm-fused", nerr(full, fused))
    assert nerr(fused, want) < 2e-2 and nerr(plain, want) < 2e-2 and nerr(full, want) < 2e-2
    assert torch.equal(fused.argmax(-1), want.argmax(-1)) and torch.equal(full.argmax(-1), want.argmax(-1))
    assert nerr(full, torch.from_numpy(g["logits"])) < 2e-2
    assert nerr(fused, torch.from_numpy(g["logits"])) < 2e-2


@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_long_sequences_up_to_64_tokens(precision, tol):
    """33..64 frames / slots per sequence take the 48 / 64-row attention tiles (the reference's position table
    allows 256 frames; its CLIs sample 16)."""
    cfg = StltModelConfig(num_classes=174, unique_categories=4, num_spatial_layers=2, num_temporal_layers=2)
    torch.manual_seed(0)
    sd = random_state_dict(Stlt(cfg).state_dict(), seed=71)
    m = _model(cfg, sd, precision)
    for frames, objects in ((39, 4), (63, 2), (16, 40)):
        batch = make_batch(3, "something", ragged=True, seed=frames, num_frames=frames, max_objects=objects)
        with torch.no_grad():
            want = O.stlt_forward(sd, batch, num_spatial_layers=2, num_temporal_layers=2)
            got = m(to_cuda(batch))["stlt"].cpu()
        assert nerr(got, want) < tol, (frames, objects, nerr(got, want))


@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_long_sequences_up_to_the_position_table(precision, tol):
    """65..256 frames (the reference's position table, models.py:88-96) take attention_long.cu: one CTA per
    (sequence, head), online softmax over 64-key blocks, causal blocks skipped. Ragged lengths, so the padded
    frames exercise the key mask inside and across key blocks; L = 256 is the maximum the reference accepts."""
    cfg = StltModelConfig(num_classes=174, unique_categories=4, num_spatial_layers=1, num_temporal_layers=2)
    torch.manual_seed(0)
    sd = random_state_dict(Stlt(cfg).state_dict(), seed=72)
    m = _model(cfg, sd, precision)
    for frames in (64, 100, 129, 255):
        batch = make_batch(3, "something", ragged=True, seed=frames, num_frames=frames, max_objects=2)
        with torch.no_grad():
            want = O.stlt_forward(sd, batch, num_spatial_layers=1, num_temporal_layers=2)
            got = m(to_cuda(batch))["stlt"].cpu()
        assert nerr(got, want) < tol, (frames, nerr(got, want))
    with pytest.raises(Exception, match="256"), torch.no_grad():
        m(to_cuda(make_batch(1, "something", num_frames=256)))


def test_random_shapes_fuzz_bf16_and_fp32():
    """Random (batch, frames, slots, layout) combinations — odd tile counts, single-frame videos, S = 1, scores on /
    off — through both precision modes (the bf16 one on the LayerNorm-fused epilogues) vs the oracle."""
    import random
    rng = random.Random(1234)
    for trial in range(14):
        layout = rng.choice(["something", "action_genome"])
        spec = stlt_b200.SOMETHING_ELSE if layout == "something" else stlt_b200.ACTION_GENOME
        ns, nt = rng.choice([(1, 1), (2, 1), (1, 3), (2, 2)])
        cfg = StltModelConfig(num_classes=spec["num_classes"], unique_categories=spec["unique_categories"],
                              num_spatial_layers=ns, num_temporal_layers=nt)
        torch.manual_seed(0)
        sd = random_state_dict(Stlt(cfg).state_dict(), seed=100 + trial)
        B = rng.choice([1, 2, 7, 26, 77, 130])
        frames = rng.choice([1, 2, 5, 16, 31])
        objects = rng.choice([0, 1, 4, 10, 31]) if frames * B < 2000 else 4
        batch = make_batch(B, layout, ragged=True, seed=trial, num_frames=frames, max_objects=objects)
        with torch.no_grad():
            want = O.stlt_forward(sd, batch, num_spatial_layers=ns, num_temporal_layers=nt)
        for precision, tol in (("bf16", BF16_TOL), ("fp32", FP32_TOL)):
            m = _model(cfg, sd, precision)
            with torch.no_grad():
                got = m(to_cuda(batch))["stlt"].cpu()
            err = nerr(got, want)
            assert torch.isfinite(got).all() and err < tol, (trial, layout, B, frames, objects, ns, nt, precision, err)


def test_two_devices_in_one_process():
    """One handle per device; kernel attributes (dynamic shared memory opt-in) are per device too."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cfg = StltModelConfig(num_classes=174, unique_categories=4, num_spatial_layers=1, num_temporal_layers=1)
    torch.manual_seed(0)
    sd = random_state_dict(Stlt(cfg).state_dict(), seed=81)
    batch = make_batch(9, "something", ragged=True, seed=82)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        for precision in ("bf16", "fp32"):
            m = Stlt(cfg, precision=precision)
            m.load_state_dict(sd)
            m = m.to(dev)
            m.train(False)
            with torch.no_grad():
                outs.append(m({k: v.to(dev) for k, v in batch.items()})["stlt"].cpu())
    assert torch.equal(outs[0], outs[2]) and torch.equal(outs[1], outs[3])


def test_backbone_forward_returns_every_frame_seq_first():
    """StltBackbone.forward (reference models.py:136-152): [L, B, H], what the reference fusion models consume."""
    cfg, sd, batch, g = golden_model_case("something")
    m = _model(cfg, sd, "fp32")
    with torch.no_grad():
        out = m.backbone(to_cuda(batch))
    B, L, _ = batch["categories"].shape
    assert tuple(out.shape) == (L, B, 768)
    for b in range(B):
        n = int(batch["lengths"][b])
        assert nerr(out[:n, b], torch.from_numpy(g["temporal"][b, :n])) < FP32_TOL
    assert len(m.state_dict()) == 174 and len(m.backbone.state_dict()) == 168  # the runner adds nothing


@pytest.mark.parametrize("layout", ["something", "action_genome"])
def test_pad_skipping_layout_on_ragged_and_degenerate_batches(layout):
    """The pad-skipping row layout (csrc/compact.cu) decides on the device which rows exist. Ragged batches, batches
    whose frames are all single-token (no objects anywhere), all-full batches and out-of-pattern inputs (an extract
    frame that carries objects, objects in padding frames) must give the logits of the padded grid / the oracle."""
    spec = stlt_b200.SOMETHING_ELSE if layout == "something" else stlt_b200.ACTION_GENOME
    cfg = StltModelConfig(num_classes=spec["num_classes"], unique_categories=spec["unique_categories"],
                          num_spatial_layers=2, num_temporal_layers=2)
    torch.manual_seed(0)
    sd = random_state_dict(Stlt(cfg).state_dict(), seed=61)
    m = _model(cfg, sd, "bf16")
    cases = {
        "ragged": make_batch(300, layout, ragged=True, seed=62),
        "dense": make_batch(130, layout, ragged=False, seed=63),
        "no_objects": make_batch(40, layout, ragged=True, seed=64, max_objects=0),
        "single_frame": make_batch(9, layout, ragged=True, seed=65, num_frames=1),
    }
    odd = make_batch(50, layout, ragged=True, seed=66)
    odd["categories"][:, :, 1] = spec["object_ids"][0]   # every frame (extract and padding frames too) carries an object
    odd["boxes"][:, :, 1] = 0.25
    if "scores" in odd:
        odd["scores"][:, :, 1] = 0.9
    cases["objects_everywhere"] = odd
    for name, batch in cases.items():
        gb = to_cuda(batch)
        with torch.no_grad():
            m.set_compaction(True)
            got = m(gb)["stlt"].float().cpu()
            m.set_compaction(False)
            padded = m(gb)["stlt"].float().cpu()
            want = O.stlt_forward(sd, batch, num_spatial_layers=2, num_temporal_layers=2)
        m.check_inputs()
        assert torch.isfinite(got).all(), name
        assert nerr(got, want) < BF16_TOL and nerr(got, padded) < 5e-3, (name, nerr(got, want), nerr(got, padded))


def test_bench_batch_parity_256_videos_both_precisions_identical_top1():
    """The batch bench.py times (BASELINE configs[1]: dense Something-Else layouts, batch 4096, weights seed 0, data seed
    100): logits of its first 256 videos, taken from forwards of the WHOLE batch on the default path (LayerNorms and
    attention in the GEMM epilogues, pad-skipping layout), against the reference's CPU forward (oracle/_ref when built,
    else the oracle port). fp32 <= 1e-4, bf16 <= 2e-2, identical arg-max on all 256 in both precisions."""
    import bench
    cfg = StltModelConfig(num_classes=174, unique_categories=4)
    torch.manual_seed(0)
    model = Stlt(cfg, precision="bf16")
    sd = random_state_dict(model.state_dict(), seed=0)
    model.load_state_dict(sd)
    model = model.cuda()
    model.train(False)
    data = make_batch(4096, "something", ragged=False, seed=100)
    host = {k: data[k] for k in ("categories", "boxes", "frame_types", "lengths")}
    dev = to_cuda(host)
    stamp = bench.parity_stamp(model, host, dev, bench.CpuReference("something", sd), torch)
    print(stamp)
    assert stamp["videos"] == 256 and stamp["ok"]
    assert stamp["fp32"] < FP32_TOL and stamp["bf16"] < BF16_TOL
    assert stamp["fp32_top1_agree"] == 256 and stamp["bf16_top1_agree"] == 256


def test_forward_cuda_graph_mode_matches_eager_and_follows_weight_updates():
    """Stlt.enable_cuda_graphs (what bench.py times): the first forward of a shape is eager, the second is captured, later
    ones replay. Replays must equal the eager logits bit for bit, for new data of the same shape, for a second shape, through
    HostPipeline, after in-place parameter updates PyTorch's version counters see (re-pack before the replay) and after
    raw `.data` writes announced with mark_weights_dirty()."""
    from stlt_b200.pipeline import HostPipeline
    cfg = StltModelConfig(num_classes=174, unique_categories=4, num_spatial_layers=2, num_temporal_layers=2)
    torch.manual_seed(0)
    sd = random_state_dict(Stlt(cfg).state_dict(), seed=71)
    m = _model(cfg, sd, "bf16")
    b1 = to_cuda(make_batch(37, "something", ragged=True, seed=72))
    b2 = to_cuda(make_batch(37, "something", ragged=True, seed=73))
    b3 = to_cuda(make_batch(5, "something", ragged=False, seed=74))
    keys = ("categories", "boxes", "frame_types", "lengths")
    with torch.no_grad():
        e1, e2, e3 = (m(b)["stlt"].clone() for b in (b1, b2, b3))
        m.enable_cuda_graphs(True)
        first = m(b1)["stlt"]           # eager
        second = m(b1)["stlt"]          # captured + replayed
        assert torch.equal(first, e1) and torch.equal(second, e1)
        assert torch.equal(m(b2)["stlt"], e2)                      # replay, new data
        assert torch.equal(m(b3)["stlt"], e3) and torch.equal(m(b3)["stlt"], e3)   # second shape: its own graph
        assert torch.equal(m(b1)["stlt"], e1)
        assert second.data_ptr() != m(b1)["stlt"].data_ptr()       # every call returns a fresh tensor
        host = [{k: b[k].cpu().pin_memory() for k in keys} for b in (b1, b2, b1)]
        got = [h.clone() for h in HostPipeline(m, "stlt").run(host)]
        assert torch.equal(got[0].cuda(), e1) and torch.equal(got[1].cuda(), e2) and torch.equal(got[2].cuda(), e1)
        # (a) an update PyTorch sees: version counter bumps -> bf16 copies re-packed eagerly, same graph replayed
        m.backbone.transformer.layers[0].linear1.weight.mul_(0.5)
        m.prediction_head.fc2.bias.add_(1.0)
        g = m(b2)["stlt"].clone()
        m.enable_cuda_graphs(False)
        want = m(b2)["stlt"].clone()
        assert torch.equal(g, want) and not torch.equal(g, e2)
        # (b) a write behind PyTorch's back: needs mark_weights_dirty()
        m.enable_cuda_graphs(True)
        m(b2), m(b2)
        m.backbone.frames_embeddings.layout_embedding.transformer.layers[1].self_attn.in_proj_weight.data.mul_(1.5)
        m.mark_weights_dirty()
        g = m(b2)["stlt"].clone()   # eager again (graphs were dropped)
        g2 = m(b2)["stlt"].clone()  # re-captured
        m.enable_cuda_graphs(False)
        want2 = m(b2)["stlt"].clone()
        assert torch.equal(g, want2) and torch.equal(g2, want2) and not torch.equal(want2, want)


def test_profile_by_role_splits_the_gemm_category_exactly():
    """stlt_get_profile_by_role (bench.py's per-kernel roofline rows): the roles partition the "gemm" category — same
    launches, FLOPs and (to rounding) time — and name what each precision mode runs: the bf16 path has the in-projection
    inside qkv_attention_kernel, the fp32-parity mode a plain in-projection GEMM; the pruned last layers run 3 GEMMs on
    the compacted rows, so every role counts one launch per layer."""
    ns, nt = 2, 3
    cfg = StltModelConfig(num_classes=174, unique_categories=4, num_spatial_layers=ns, num_temporal_layers=nt)
    torch.manual_seed(0)
    sd = random_state_dict(Stlt(cfg).state_dict(), seed=81)
    batch = to_cuda(make_batch(19, "something", ragged=True, seed=82))
    for precision, in_role in (("bf16", "qkv_attention"), ("fp32", "in_proj")):
        m = _model(cfg, sd, precision)
        with torch.no_grad():
            m(batch)
            m.set_profiling(True)
            m(batch)
            prof = m.get_profile()
            roles = m.get_profile_by_role()
            m.set_profiling(False)
        assert set(roles) == {in_role, "out_proj", "linear1", "linear2"}, (precision, roles)
        assert all(v["launches"] == ns + nt for v in roles.values()), (precision, roles)
        assert sum(v["launches"] for v in roles.values()) == prof["gemm"]["launches"]
        assert sum(v["flops"] for v in roles.values()) == pytest.approx(prof["gemm"]["flops"], rel=1e-12)
        assert sum(v["ms"] for v in roles.values()) == pytest.approx(prof["gemm"]["ms"], rel=1e-6)
        assert all(v["ms"] > 0 and v["flops"] > 0 for v in roles.values())
        # a profile that was not refreshed is not re-reported: the next get_profile() without spans clears the roles
        m.get_profile()
        assert m.get_profile_by_role() == {}


@pytest.mark.parametrize("layout", ["something", "action_genome"])
def test_fp32_parity_mode_with_fused_layer_norm_matches_separate_kernels_and_oracle(layout):
    """The fp32-parity mode folds its LayerNorms into the epilogues of the 3-term GEMMs (split operands, split folded
    weights, hi / lo copy of the residual stream as a by-product). A/B against the separate add + LayerNorm kernels,
    on the pad-skipping layout and on the padded grid, all inside the 1e-4 gate of the reference golden."""
    from tests.util import golden_model_case
    cfg, sd, batch, g = golden_model_case(layout)
    model = _model(cfg, sd, "fp32")
    gb = to_cuda(batch)
    want = torch.from_numpy(g["logits"])
    got, launches = {}, {}
    with torch.no_grad():
        for name, (ln, cp) in {"fused": (True, True), "fused_padded_grid": (True, False), "separate": (False, True),
                               "separate_padded_grid": (False, False)}.items():
            model.set_fused_layer_norm(ln)
            model.set_compaction(cp)
            got[name] = model(gb)["stlt"].float().cpu()
            launches[name] = model.last_launch_count()
    print(layout, {k: nerr(v, want) for k, v in got.items()}, launches)
    for name, v in got.items():
        assert nerr(v, want) < FP32_TOL, (name, nerr(v, want))
        assert torch.equal(v.argmax(-1), want.argmax(-1))
    assert nerr(got["fused"], got["separate"]) < 5e-5
    # 24 residual + LayerNorm launches gone, one LayerNorm of the pooled rows added (as in bf16 mode)
    assert launches["fused_padded_grid"] == launches["separate_padded_grid"] - 23
    assert launches["fused"] == launches["separate"] - 23
