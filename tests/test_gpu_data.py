"""GPU parity tests of the §8(f) rows built so far: the device batch builder (StltDataset.__getitem__
+ StltCollater) and the top-k evaluator counters."""
import json

import numpy as np
import pytest
import torch

import stlt_b200
from oracle import stlt_oracle as O
from oracle.make_golden import synth_dataset_json
from stlt_b200 import LayoutStore, Stlt, StltModelConfig, TopKCounter
from stlt_b200.synthetic import random_state_dict
from tests.util import load_golden, nerr

pytestmark = pytest.mark.gpu


def _assert_batch_equal(got, want, keys):
    for k in keys:
        g, w = got[k].cpu().numpy(), want[k].numpy() if isinstance(want[k], torch.Tensor) else want[k]
        if g.dtype == np.float32:
            assert np.array_equal(g.view(np.uint32), w.view(np.uint32)), k
        else:
            assert np.array_equal(g, w), k


@pytest.mark.parametrize("dataset", ["something", "action_genome"])
def test_build_batch_matches_reference_collater_golden(dataset):
    """JSON -> CSR store -> one kernel == the reference StltDataset + StltCollater, bit for bit."""
    g = load_golden(f"collate_{dataset}.npz")
    meta = json.loads(bytes(g["json"]).decode())
    store = LayoutStore(dataset, meta["videos"], meta["sizes"])
    assert store.max_num_objects == int(g["max_num_objects"])
    batch = store.build_batch(list(range(len(store))), check=True)
    keys = ["categories", "boxes", "frame_types", "lengths", "src_key_padding_mask_boxes", "src_key_padding_mask_frames"]
    if dataset == "action_genome":
        keys.append("scores")
    else:
        assert "scores" not in batch
    _assert_batch_equal(batch, g, keys)


@pytest.mark.parametrize("dataset", ["something", "action_genome"])
def test_build_batch_matches_oracle_on_larger_dataset_and_train_indices(dataset):
    videos, _, sizes = synth_dataset_json(dataset, n_videos=120, seed=77)
    store = LayoutStore(dataset, videos, sizes)
    mno = store.max_num_objects
    order = [5, 0, 119, 37, 64, 64, 3]
    # test-time sampling
    want = O.collate([O.build_sample(videos[i], *sizes[videos[i]["id"]], dataset, mno) for i in order], dataset, mno)
    got = store.build_batch(order, check=True)
    keys = ["categories", "boxes", "frame_types", "lengths", "src_key_padding_mask_boxes", "src_key_padding_mask_frames"]
    keys += ["scores"] if dataset == "action_genome" else []
    _assert_batch_equal(got, want, keys)
    # explicit (training-style) frame indices
    rng = np.random.default_rng(3)
    idx = [sorted(rng.integers(0, len(videos[i]["frames"]), size=min(16, len(videos[i]["frames"]))).tolist()) for i in order]
    want = O.collate([O.build_sample(videos[i], *sizes[videos[i]["id"]], dataset, mno, indices=ix)
                      for i, ix in zip(order, idx)], dataset, mno)
    got = store.build_batch(order, frame_indices=idx, check=True)
    _assert_batch_equal(got, want, keys)
    # whole dataset in one batch, empty batch
    everything = store.build_batch(list(range(len(store))), check=True)
    assert everything["categories"].shape[0] == 120
    assert store.build_batch([])["categories"].shape[0] == 0


def test_json_to_logits_end_to_end():
    """Raw layouts -> device batch builder -> CUDA forward == oracle data path -> oracle forward."""
    dataset = "something"
    videos, _, sizes = synth_dataset_json(dataset, n_videos=12, seed=5)
    store = LayoutStore(dataset, videos, sizes)
    mno = store.max_num_objects
    cfg = StltModelConfig(num_classes=174, unique_categories=4)
    torch.manual_seed(0)
    sd = random_state_dict(Stlt(cfg).state_dict(), seed=71)
    model = Stlt(cfg)
    model.load_state_dict(sd)
    model = model.to("cuda")
    model.train(False)
    batch = store.build_batch(list(range(12)))
    with torch.no_grad():
        got = model(batch)["stlt"].cpu()
        ref_batch = O.collate([O.build_sample(v, *sizes[v["id"]], dataset, mno) for v in videos], dataset, mno)
        want = O.stlt_forward(sd, ref_batch)
    assert nerr(got, want) < 1e-4


def test_topk_counter_matches_reference_evaluator_semantics():
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(1000, 174, generator=g)
    labels = torch.randint(0, 174, (1000,), generator=g)
    # exact ties: argmax picks the lowest index (matched); torch.topk's tie order is unspecified, so
    # the top-5 check uses tie-free rows only
    logits[7, :3] = logits[7].max() + 1.0
    labels[7] = 0
    top1 = (logits.argmax(-1) == labels).sum().item()                                   # evaluation.py:24-26
    top5 = (logits.topk(k=5).indices == labels.unsqueeze(1)).any(dim=1).sum().item()    # evaluation.py:27-34
    counter = TopKCounter(total_instances=1000)
    counter.process(logits[:600].cuda(), labels[:600].cuda())
    counter.process(logits[600:].cuda(), labels[600:].cuda())
    m = counter.evaluate()
    assert m["stlt_top1_accuracy"] == top1 / 1000 and m["stlt_top5_accuracy"] == top5 / 1000
    counter.reset()
    assert counter.evaluate()["stlt_top1_accuracy"] == 0


def test_charades_map_evaluator_matches_reference_golden():
    """Device EvaluatorActionGenome: sigmoid + label accumulation over ragged batches, mAP on the device, vs
    the reference's charades_map (tests/golden/charades_map.npz) and the oracle."""
    import numpy as np
    from oracle import stlt_oracle
    from stlt_b200 import CharadesMapEvaluator
    from tests.util import load_golden
    g = load_golden("charades_map.npz")
    logits, n = torch.from_numpy(g["logits"]).cuda(), g["logits"].shape[0]
    for labels_key, map_key, aps_key in (("labels2", "map2", "aps2"), ("labels", "map", "aps")):
        labels = torch.from_numpy(g[labels_key]).cuda()
        ev = CharadesMapEvaluator(n, logits.shape[1])
        i = 0
        for size in (100, 1, 333, 343):
            ev.process({"stlt": logits[i:i + size]}, labels[i:i + size])
            i += size
        assert i == n
        got = ev.evaluate()["map"]
        aps = ev.average_precisions.cpu().numpy()
        want, want_aps = float(g[map_key]), g[aps_key]
        assert np.allclose(aps, want_aps, rtol=1e-6, atol=0, equal_nan=True), np.nanmax(np.abs(aps - want_aps))
        assert (np.isnan(got) and np.isnan(want)) or abs(got - want) < 1e-6
        assert torch.equal(ev.ground_truths, labels)
        assert float((ev.predictions - torch.sigmoid(logits)).abs().max()) < 2e-7
    assert ev.is_best() is False  # nan mAP never counts as an improvement (nan > 0 is False), as in the reference
