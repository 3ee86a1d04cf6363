"""CACNF on precomputed appearance features. CPU: the oracle restatement vs
tests/golden/cacnf_something.npz (the unmodified reference CrossAttentionCentralNetFusion, eval mode,
ResNet features injected) and the drop-in module's checkpoint contract. GPU: the CUDA path vs both."""
import pytest
import torch

from oracle import stlt_oracle
from tests.util import load_golden, nerr, to_cuda, weights_checksum

NAMES = ("stlt", "resnet3d", "caf", "ensemble")


def _case():
    import stlt_b200
    from stlt_b200.synthetic import make_appearance_features, make_batch, random_state_dict
    g = load_golden("cacnf_something.npz")
    cfg = stlt_b200.CacnfModelConfig(num_classes=174, unique_categories=4)
    torch.manual_seed(0)
    model = stlt_b200.Cacnf(cfg)
    sd = random_state_dict(model.state_dict(), seed=int(g["weight_seed"]))
    assert len(sd) == int(g["num_entries"]) == 360
    assert abs(weights_checksum(sd) - float(g["weights_checksum"])) < 1e-6 * float(g["weights_checksum"]), \
        "seeded CACNF weights differ from the ones the reference golden was generated with (names / order / shapes)"
    batch = make_batch(int(g["batch_size"]), layout="something", ragged=True, seed=int(g["batch_seed"]))
    feats = make_appearance_features(int(g["batch_size"]), seed=int(g["batch_seed"]) + 50)
    return cfg, model, sd, batch, feats, g


def test_oracle_cacnf_matches_reference_golden():
    cfg, model, sd, batch, feats, g = _case()
    with torch.no_grad():
        out = stlt_oracle.cacnf_forward(sd, batch, feats)
    for name in NAMES:
        want = torch.from_numpy(g["logits_" + name])
        assert nerr(out[name], want) < 2e-5, name


def test_cacnf_module_contract():
    cfg, model, sd, batch, feats, g = _case()
    assert model.logit_names == NAMES
    model.load_reference_state_dict({**sd, "backbone.appearance_branch.resnet.resnet.0.weight": torch.zeros(3)})
    keys = list(model.state_dict().keys())
    assert "backbone.appearance_branch.projector.weight" in keys
    assert tuple(model.state_dict()["backbone.appearance_branch.projector.weight"].shape) == (768, 2048, 1, 1, 1)
    assert tuple(model.state_dict()["backbone.appearance_branch.pos_embed"].shape) == (33, 1, 768)
    assert tuple(model.state_dict()["fusion_classifier.fc1.weight"].shape) == (768, 1536)
    assert sum(k.startswith("backbone.mm_fusion.") for k in keys) == 4 * 30
    with pytest.raises(RuntimeError, match="no CPU path"):
        model.train(False)
        with torch.no_grad():
            model({**batch, "video_features": feats})


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("bf16", 2e-2), ("fp32", 1e-4)])
def test_cacnf_gpu_matches_oracle_and_golden(precision, tol):
    cfg, model, sd, batch, feats, g = _case()
    model.load_state_dict(sd)
    model.precision = precision
    model = model.cuda()
    model.train(False)
    with torch.no_grad():
        got = model({**to_cuda(batch), "video_features": feats.cuda()})
        want = stlt_oracle.cacnf_forward(sd, batch, feats)
    for name in NAMES:
        e1 = nerr(got[name], want[name])
        e2 = nerr(got[name], torch.from_numpy(g["logits_" + name]))
        print(name, e1, e2)
        assert torch.isfinite(got[name]).all()
        assert e1 < tol and e2 < tol, (name, e1, e2)
    mean = ((got["stlt"].double() + got["resnet3d"].double()) + got["caf"].double()) / 3.0
    assert float((got["ensemble"].double() - mean).abs().max()) < 1e-6
    top_ref = torch.from_numpy(g["logits_ensemble"]).argmax(-1)
    assert torch.equal(got["ensemble"].argmax(-1).cpu(), top_ref)


@pytest.mark.gpu
@pytest.mark.parametrize("Tq,Tk,causal,masked", [(33, 33, False, False), (17, 33, False, False), (33, 17, False, True),
                                                 (17, 17, True, True), (64, 64, True, True), (1, 40, False, True),
                                                 (48, 5, False, False)])
def test_attention_cross_kernel(Tq, Tk, causal, masked):
    import ctypes
    from stlt_b200 import lib as L
    lib = L.load_library()
    dims = L.StltDims(768, 12, 0, 0, 4, 174, 256, 5, 1e-12, 1e-5)
    h = ctypes.c_void_p()
    L.check(None, lib.stlt_create(ctypes.byref(dims), ctypes.byref(h)))
    N = 37
    g = torch.Generator(device="cuda").manual_seed(Tq * 100 + Tk)
    qsrc = torch.randn(N * Tq, 2304, device="cuda", generator=g).to(torch.bfloat16)
    kvsrc = torch.randn(N * Tk, 2304, device="cuda", generator=g).to(torch.bfloat16)
    mask_src = None
    if masked:
        mask_src = (torch.rand(N, Tk, device="cuda", generator=g) > 0.3).long()
        mask_src[:, 0] = 1
    out = torch.full((N * Tq, 768), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(h, lib.stlt_op_attention_cross(h, torch.cuda.current_stream().cuda_stream, qsrc.data_ptr(), kvsrc.data_ptr(),
                                           mask_src.data_ptr() if masked else None, N, Tq, Tk, int(causal), out.data_ptr()))
    q = qsrc.float()[:, :768].view(N, Tq, 12, 64).transpose(1, 2)
    k = kvsrc.float()[:, 768:1536].view(N, Tk, 12, 64).transpose(1, 2)
    v = kvsrc.float()[:, 1536:].view(N, Tk, 12, 64).transpose(1, 2)
    scores = q @ k.transpose(-1, -2) / 8.0
    if masked:
        scores = scores.masked_fill((mask_src == 0).view(N, 1, 1, Tk), float("-inf"))
    if causal:
        scores = scores.masked_fill(torch.triu(torch.ones(Tq, Tk, dtype=torch.bool, device="cuda"), diagonal=1), float("-inf"))
    want = (torch.softmax(scores, -1) @ v).transpose(1, 2).reshape(N * Tq, 768)
    assert nerr(out, want) < 1e-2
    lib.stlt_destroy(h)


def _variant_case(name):
    import stlt_b200
    from stlt_b200.synthetic import make_appearance_features, make_batch, random_state_dict
    g = load_golden("caf_lcf_something.npz")
    cfg = stlt_b200.CacnfModelConfig(num_classes=174, unique_categories=4)
    torch.manual_seed(0)
    model = (stlt_b200.Caf if name == "caf" else stlt_b200.Lcf)(cfg)
    sd = random_state_dict(model.state_dict(), seed=int(g["weight_seed"]))
    assert len(sd) == int(g[f"entries_{name}"])
    assert abs(weights_checksum(sd) - float(g[f"checksum_{name}"])) < 1e-6 * float(g[f"checksum_{name}"])
    batch = make_batch(int(g["batch_size"]), layout="something", ragged=True, seed=int(g["batch_seed"]))
    feats = make_appearance_features(int(g["batch_size"]), seed=int(g["batch_seed"]) + 50)
    return model, sd, batch, feats, torch.from_numpy(g[f"logits_{name}"])


@pytest.mark.parametrize("name", ["caf", "lcf"])
def test_oracle_caf_lcf_match_reference_golden(name):
    model, sd, batch, feats, want = _variant_case(name)
    fn = stlt_oracle.caf_forward if name == "caf" else stlt_oracle.lcf_forward
    with torch.no_grad():
        got = fn(dict(sd), batch, feats)[name]
    assert nerr(got, want) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["caf", "lcf"])
@pytest.mark.parametrize("precision,tol", [("bf16", 2e-2), ("fp32", 1e-4)])
def test_caf_lcf_gpu_match_reference_golden(name, precision, tol):
    model, sd, batch, feats, want = _variant_case(name)
    model.load_state_dict(sd)
    model.precision = precision
    model = model.cuda()
    model.train(False)
    with torch.no_grad():
        out = model({**to_cuda(batch), "video_features": feats.cuda()})
    assert list(out) == [name]
    err = nerr(out[name], want)
    print(name, precision, err)
    assert err < tol
