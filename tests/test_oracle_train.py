"""CPU: the oracle's restatement of the training step (loss, gradients, clip + AdamW + schedule)
against tests/golden/train_*.npz, which oracle/make_golden.py produced by running the UNMODIFIED
reference loop body (src/train.py:117-135) with dropout p = 0."""
import numpy as np
import pytest
import torch

from oracle import stlt_oracle
from tests.util import load_golden, weights_checksum


def _case(layout):
    import stlt_b200
    from stlt_b200.synthetic import make_batch, random_state_dict
    g = load_golden(f"train_{layout}.npz")
    spec = stlt_b200.SOMETHING_ELSE if layout == "something" else stlt_b200.ACTION_GENOME
    cfg = stlt_b200.StltModelConfig(num_classes=spec["num_classes"], unique_categories=spec["unique_categories"],
                                    hidden_dropout_prob=0.0)
    torch.manual_seed(0)
    shapes = stlt_b200.Stlt(cfg).state_dict()
    sd = random_state_dict(shapes, seed=int(g["weight_seed"]))
    assert abs(weights_checksum(sd) - float(g["weights_checksum"])) < 1e-6 * float(g["weights_checksum"])
    batch = make_batch(int(g["batch_size"]), layout=layout, ragged=True, seed=int(g["batch_seed"]))
    labels = torch.from_numpy(g["labels"])
    loss = "cross_entropy" if layout == "something" else "bce_with_logits"
    return cfg, sd, batch, labels, loss, g


@pytest.mark.parametrize("layout", ["something", "action_genome"])
def test_oracle_gradients_match_reference(layout):
    cfg, sd, batch, labels, loss, g = _case(layout)
    value, logits, grads = stlt_oracle.loss_and_grads(sd, batch, labels, loss)
    assert abs(float(value) - float(g["losses"][0])) < 2e-6 * max(1.0, abs(float(g["losses"][0])))
    assert np.abs(logits.numpy() - g["logits0"]).max() < 2e-5 * np.abs(g["logits0"]).max()
    names = [str(n) for n in g["param_names"]]
    for name, want in zip(names, g["grad_norms"]):
        if want < 0:  # grad is None in the reference (orphan layer / unused score embedding)
            assert name not in grads, name
            continue
        got = float(grads[name].norm())
        assert abs(got - want) <= 2e-4 * want + 1e-9, (name, got, want)
    params = {k: v for k, v in grads.items()}
    for key in g:
        if not key.startswith("grad/"):
            continue
        want = torch.from_numpy(g[key])
        name = key[5:]
        if name == "position_rows":
            got = params["backbone.frames_embeddings.position_embeddings.weight"][:17]
        elif name == "sp0_in_proj_rows":
            got = params["backbone.frames_embeddings.layout_embedding.transformer.layers.0.self_attn.in_proj_weight"][::96, ::8]
        elif name == "tm7_linear2_rows":
            got = params["backbone.transformer.layers.7.linear2.weight"][::32, ::64]
        else:
            got = params[name]
        err = float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))
        assert err < 5e-4, (key, err)


@pytest.mark.parametrize("layout", ["something", "action_genome"])
def test_oracle_adamw_steps_match_reference(layout):
    cfg, sd, batch, labels, loss, g = _case(layout)
    sd = {k: v.clone() for k, v in sd.items()}
    names = [str(n) for n in g["param_names"]]
    state = {}
    for step in range(1, int(g["steps"]) + 1):
        value, _, grads = stlt_oracle.loss_and_grads(sd, batch, labels, loss)
        lr = 5e-5 * stlt_oracle.linear_schedule(step - 1, 2, 10)
        total = stlt_oracle.adamw_update(sd, grads, state, step, lr)
        assert abs(float(value) - float(g["losses"][step - 1])) < 2e-3 * abs(float(g["losses"][step - 1]))
        assert abs(total - float(g["total_grad_norms"][step - 1])) < 2e-3 * float(g["total_grad_norms"][step - 1])
        want_norms = g[f"param_norms_step{step}"]
        want_sums = g[f"param_sums_step{step}"]
        for name, wn, ws in zip(names, want_norms, want_sums):
            p = sd[name].double()
            assert abs(float(p.norm()) - wn) <= 1e-5 * wn + 1e-7, (step, name)
            assert abs(float(p.sum()) - ws) <= 1e-4 * max(abs(ws), wn) + 1e-6, (step, name)


def test_criterion_matches_torch():
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(7, 174, generator=g)
    labels = torch.randint(0, 174, (7,), generator=g)
    assert torch.allclose(stlt_oracle.criterion(logits, labels, "cross_entropy"),
                          torch.nn.functional.cross_entropy(logits, labels), atol=1e-6)
    targets = (torch.rand(7, 157, generator=g) < 0.1).float()
    x = torch.randn(7, 157, generator=g) * 4
    assert torch.allclose(stlt_oracle.criterion(x, targets, "bce_with_logits"),
                          torch.nn.functional.binary_cross_entropy_with_logits(x, targets), atol=1e-6)
