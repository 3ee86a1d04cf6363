"""GPU parity tests of the kernels the bf16 headline number runs on, called one at a time through the C ABI:
the LayerNorm-fused GEMM epilogues (GEMM_EPI_NORM_A / GEMM_EPI_RESID, reference math
src/modelling/models.py:46-55 via nn.TransformerEncoderLayer) and the in-projection + attention kernel
(gemm_qkv_attn.cu). References are fp64 torch on the same bf16-rounded operands."""
import ctypes
import math

import pytest
import torch

from stlt_b200 import lib as L
from tests.util import nerr

pytestmark = pytest.mark.gpu

H = 768


@pytest.fixture(scope="module")
def handle():
    lib = L.load_library()
    dims = L.StltDims(768, 12, 0, 0, 4, 174, 256, 5, 1e-12, 1e-5)
    h = ctypes.c_void_p()
    L.check(None, lib.stlt_create(ctypes.byref(dims), ctypes.byref(h)))
    yield h
    lib.stlt_destroy(h)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _row_stats(z: torch.Tensor) -> torch.Tensor:
    """Per-row (sum, sum of squares) of z split over the 6 slots the producing GEMM writes (128-column slabs)."""
    zs = z.double().view(z.shape[0], 6, 128)
    return torch.stack([zs.sum(-1), (zs * zs).sum(-1)], dim=-1).float().contiguous()  # [M, 6, 2]


def _layer_norm64(z: torch.Tensor, gamma, beta, eps) -> torch.Tensor:
    z = z.double()
    mu = z.mean(-1, keepdim=True)
    var = z.var(-1, unbiased=False, keepdim=True)
    return (z - mu) / torch.sqrt(var + eps) * gamma.double() + beta.double()


def _pack_folded(handle, w, gamma, beta, bias, head_major=0):
    lib = L.load_library()
    n, k = w.shape
    wf = torch.empty(n, k, dtype=torch.bfloat16, device="cuda")
    s = torch.empty(n, device="cuda")
    c = torch.empty(n, device="cuda")
    L.check(handle, lib.stlt_op_pack_folded(
        handle, _stream(), w.data_ptr(), gamma.data_ptr() if gamma is not None else None,
        beta.data_ptr() if beta is not None else None, bias.data_ptr(), n, k, wf.data_ptr(), s.data_ptr(),
        c.data_ptr(), head_major))
    return wf, s, c


def _head_major_rows():
    idx = torch.arange(2304)
    return ((idx % 192) // 64) * 768 + (idx // 192) * 64 + idx % 64


def test_pack_folded_head_major(handle):
    g = torch.Generator(device="cuda").manual_seed(5)
    w = torch.randn(2304, H, device="cuda", generator=g) / math.sqrt(H)
    gamma = 1 + 0.1 * torch.randn(H, device="cuda", generator=g)
    beta = 0.1 * torch.randn(H, device="cuda", generator=g)
    bias = torch.randn(2304, device="cuda", generator=g)
    wf, s, c = _pack_folded(handle, w, gamma, beta, bias)
    wh, sh, ch = _pack_folded(handle, w, gamma, beta, bias, head_major=1)
    perm = _head_major_rows().cuda()
    assert torch.equal(wf, (w * gamma).to(torch.bfloat16))
    assert torch.equal(wh, wf[perm]) and torch.equal(sh, s[perm]) and torch.equal(ch, c[perm])
    assert nerr(c, w.double() @ beta.double() + bias.double()) < 1e-6
    # identity fold
    w0, s0, c0 = _pack_folded(handle, w, None, None, bias, head_major=1)
    assert torch.equal(w0, w.to(torch.bfloat16)[perm]) and torch.equal(c0, bias[perm])


# (rows, n, gelu): odd and even tile counts, more than two tiles per CTA pair
NORM_A_CASES = [(128, 2304, 0), (384, 3072, 2), (148 * 128 + 128, 768 * 3, 0), (640, 3072, 2)]


@pytest.mark.parametrize("m,n,gelu", NORM_A_CASES)
def test_gemm_norm_a_epilogue(handle, m, n, gelu):
    """out = act(LN(z) W^T + b) from the un-normalised bf16 stream and gamma-folded weights."""
    lib = L.load_library()
    g = torch.Generator(device="cuda").manual_seed(m + n)
    z = torch.randn(m, H, device="cuda", generator=g) * 1.7 + 0.3
    w = torch.randn(n, H, device="cuda", generator=g) / math.sqrt(H)
    gamma = 1 + 0.1 * torch.randn(H, device="cuda", generator=g)
    beta = 0.1 * torch.randn(H, device="cuda", generator=g)
    bias = 0.1 * torch.randn(n, device="cuda", generator=g)
    eps = 1e-5
    wf, s, c = _pack_folded(handle, w, gamma, beta, bias)
    zb = z.to(torch.bfloat16).contiguous()
    stats = _row_stats(z)
    out = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device="cuda")
    L.check(handle, lib.stlt_op_gemm_fused(handle, _stream(), L.GEMM_EPI_NORM_A, zb.data_ptr(), m, wf.data_ptr(), n, H,
                                           None, out.data_ptr(), None, gelu, stats.data_ptr(), s.data_ptr(),
                                           c.data_ptr(), None, eps, 1))
    torch.cuda.synchronize()
    # the kernel normalises the bf16-rounded stream with the statistics of the fp32 one
    mu = z.double().mean(-1, keepdim=True)
    rstd = 1.0 / torch.sqrt(z.double().var(-1, unbiased=False, keepdim=True) + eps)
    ref = ((zb.double() - mu) * rstd) @ wf.double().T + (w.double() @ beta.double() + bias.double())
    exact = _layer_norm64(z, gamma, beta, eps) @ w.double().T + bias.double()
    if gelu:
        ref = torch.nn.functional.gelu(ref)
        exact = torch.nn.functional.gelu(exact)
    assert torch.isfinite(out.float()).all()
    assert nerr(out.float(), ref) < 4e-3      # bf16 rounding of the output
    assert nerr(out.float(), exact) < 1.5e-2  # + bf16 rounding of the operands


@pytest.mark.parametrize("m,k,prev_norm", [(128, 768, 0), (384, 768, 1), (148 * 128 + 128, 768, 1), (256, 3072, 1),
                                           (640, 3072, 0)])
def test_gemm_resid_epilogue(handle, m, k, prev_norm):
    """z <- (LN(z) | z) + A W^T + b in place, bf16 copy and partial row statistics as by-products."""
    lib = L.load_library()
    g = torch.Generator(device="cuda").manual_seed(m + k + prev_norm)
    z = torch.randn(m, H, device="cuda", generator=g) * 1.3 - 0.2
    a = torch.randn(m, k, device="cuda", generator=g).to(torch.bfloat16).contiguous()
    w = (torch.randn(H, k, device="cuda", generator=g) / math.sqrt(k)).to(torch.bfloat16).contiguous()
    bias = 0.1 * torch.randn(H, device="cuda", generator=g)
    gamma = 1 + 0.1 * torch.randn(H, device="cuda", generator=g)
    beta = 0.1 * torch.randn(H, device="cuda", generator=g)
    eps = 1e-5
    stats_in = _row_stats(z)
    z_io = z.clone()
    zb = torch.full((m, H), float("nan"), dtype=torch.bfloat16, device="cuda")
    stats_out = torch.full((m, 6, 2), float("nan"), device="cuda")
    L.check(handle, lib.stlt_op_gemm_fused(handle, _stream(), L.GEMM_EPI_RESID, a.data_ptr(), m, w.data_ptr(), H, k,
                                           bias.data_ptr(), z_io.data_ptr(), zb.data_ptr(), 0, stats_in.data_ptr(),
                                           gamma.data_ptr(), beta.data_ptr(), stats_out.data_ptr(), eps, prev_norm))
    torch.cuda.synchronize()
    x = _layer_norm64(z, gamma, beta, eps) if prev_norm else z.double()
    ref = x + a.double() @ w.double().T + bias.double()
    assert nerr(z_io, ref) < 5e-6
    assert torch.equal(zb, z_io.to(torch.bfloat16))
    so = stats_out.double().sum(1)
    assert nerr(so[:, 0], ref.sum(-1)) < 1e-5
    assert nerr(so[:, 1], (ref * ref).sum(-1)) < 1e-5


@pytest.mark.parametrize("m,k,prev_norm", [(128, 768, 0), (384, 768, 1), (148 * 128 + 128, 768, 1), (256, 3072, 1)])
def test_gemm_resid_epilogue_hilo_planes(handle, m, k, prev_norm):
    """The same residual update on a stream stored as two bf16 planes (z = hi + lo), in place."""
    lib = L.load_library()
    g = torch.Generator(device="cuda").manual_seed(7 * m + k + prev_norm)
    z = torch.randn(m, H, device="cuda", generator=g) * 1.3 - 0.2
    hi = z.to(torch.bfloat16)
    lo = (z - hi.float()).to(torch.bfloat16)
    z_in = hi.double() + lo.double()          # what the kernel reads
    a = torch.randn(m, k, device="cuda", generator=g).to(torch.bfloat16).contiguous()
    w = (torch.randn(H, k, device="cuda", generator=g) / math.sqrt(k)).to(torch.bfloat16).contiguous()
    bias = 0.1 * torch.randn(H, device="cuda", generator=g)
    gamma = 1 + 0.1 * torch.randn(H, device="cuda", generator=g)
    beta = 0.1 * torch.randn(H, device="cuda", generator=g)
    eps = 1e-5
    stats_in = _row_stats(z_in.float())
    hi_io, lo_io = hi.clone().contiguous(), lo.clone().contiguous()
    stats_out = torch.full((m, 6, 2), float("nan"), device="cuda")
    L.check(handle, lib.stlt_op_gemm_resid_hilo(handle, _stream(), a.data_ptr(), m, w.data_ptr(), k, bias.data_ptr(),
                                                hi_io.data_ptr(), lo_io.data_ptr(), stats_in.data_ptr(), gamma.data_ptr(),
                                                beta.data_ptr(), stats_out.data_ptr(), eps, prev_norm))
    torch.cuda.synchronize()
    if prev_norm:
        st = stats_in.double().sum(1)
        mu = (st[:, 0] / H).unsqueeze(1)
        var = (st[:, 1] / H).unsqueeze(1) - mu * mu
        x = (z_in - mu) / torch.sqrt(var + eps) * gamma.double() + beta.double()
    else:
        x = z_in
    ref = x + a.double() @ w.double().T + bias.double()
    got = hi_io.double() + lo_io.double()
    assert nerr(got, ref) < 2e-5                      # two bf16 planes carry ~2^-17 relative
    assert torch.equal(hi_io, got.float().to(torch.bfloat16)) or nerr(hi_io.float(), ref) < 4e-3
    assert (lo_io.float().abs() <= hi_io.float().abs() * 2.0 ** -8 + 1e-30).all()
    so = stats_out.double().sum(1)
    assert nerr(so[:, 0], ref.sum(-1)) < 1e-5 and nerr(so[:, 1], (ref * ref).sum(-1)) < 1e-5


def _attention_ref(qkv: torch.Tensor, valid: torch.Tensor, T: int, causal: bool) -> torch.Tensor:
    """qkv fp64 [N*T, 2304] (reference row order Q | K | V), valid bool [N*T] -> ctx fp64 [N*T, 768]."""
    n = qkv.shape[0] // T
    q, k, v = (qkv[:, i * H:(i + 1) * H].view(n, T, 12, 64).transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-1, -2) / 8.0
    mask = ~valid.view(n, 1, 1, T)
    if causal:
        mask = mask | torch.ones(T, T, dtype=torch.bool, device=qkv.device).triu(1)
    p = torch.softmax(s.masked_fill(mask, float("-inf")), dim=-1)
    return (p @ v).transpose(1, 2).reshape(n * T, H)


# (sequence length, sequences, causal, deferred LayerNorm)
QKV_ATTN_CASES = [
    (5, 25, False, False),       # one row block, Something-Else slots
    (5, 25 * 2 * 74 + 3, False, True),   # every CTA pair busy + a ragged tail block (odd block count)
    (17, 7 * 5 + 1, True, True),    # temporal shape: causal + key padding, band of 48 keys
    (11, 60, False, True),       # Action-Genome slots
    (1, 300, False, False), (2, 200, True, True), (8, 64, False, True), (9, 30, True, False),
    (16, 40, True, True), (23, 21, True, True), (32, 13, False, True), (31, 9, True, False),
]


@pytest.mark.parametrize("T,n_seq,causal,prev_norm", QKV_ATTN_CASES)
def test_qkv_attention_kernel(handle, T, n_seq, causal, prev_norm):
    lib = L.load_library()
    g = torch.Generator(device="cuda").manual_seed(T * 1000 + n_seq)
    tokens = T * n_seq
    m = (tokens + 127) // 128 * 128
    z = torch.randn(m, H, device="cuda", generator=g) * (1.5 if prev_norm else 1.0) + (0.2 if prev_norm else 0.0)
    w = torch.randn(2304, H, device="cuda", generator=g) / math.sqrt(H)
    w[:1536] *= 2.0  # sharper softmax than default init: the masks matter
    bias = 0.2 * torch.randn(2304, device="cuda", generator=g)
    gamma = 1 + 0.1 * torch.randn(H, device="cuda", generator=g)
    beta = 0.1 * torch.randn(H, device="cuda", generator=g)
    eps = 1e-5
    # key padding: position 0 of every sequence is always valid (CLS slot / first frame), the rest is random
    valid = torch.rand(n_seq, T, device="cuda", generator=g) < 0.7
    valid[:, 0] = True
    mask_src = valid.view(-1).to(torch.int64).contiguous()
    zb = z.to(torch.bfloat16).contiguous()
    wh, sh, ch = _pack_folded(handle, w, gamma if prev_norm else None, beta if prev_norm else None, bias, head_major=1)
    stats = _row_stats(z)
    ctx = torch.full((m, H), float("nan"), dtype=torch.bfloat16, device="cuda")
    L.check(handle, lib.stlt_op_qkv_attention(
        handle, _stream(), zb.data_ptr(), m, tokens, wh.data_ptr(), sh.data_ptr(), ch.data_ptr(),
        stats.data_ptr() if prev_norm else None, eps, mask_src.data_ptr(), n_seq, T, int(causal), ctx.data_ptr()))
    torch.cuda.synchronize()
    # reference on the operands the kernel sees: bf16 stream, bf16 (folded) weights, bf16-rounded q / k / v
    if prev_norm:
        mu = z.double().mean(-1, keepdim=True)
        rstd = 1.0 / torch.sqrt(z.double().var(-1, unbiased=False, keepdim=True) + eps)
        wf = (w * gamma).to(torch.bfloat16)
        qkv = ((zb.double() - mu) * rstd) @ wf.double().T + (w.double() @ beta.double() + bias.double())
    else:
        qkv = zb.double() @ w.to(torch.bfloat16).double().T + bias.double()
    qkv = qkv[:tokens].to(torch.bfloat16).double()
    ref = _attention_ref(qkv, valid.view(-1), T, causal)
    got = ctx[:tokens].float()
    assert torch.isfinite(got).all()
    assert nerr(got, ref) < 1e-2  # bf16 probabilities and output
    # and against the unfused kernels of the same library on the same inputs (in-projection GEMM -> attention)
    wf_rows = (w * gamma).to(torch.bfloat16) if prev_norm else w.to(torch.bfloat16)
    qkv_dev = torch.empty(m, 2304, dtype=torch.bfloat16, device="cuda")
    if prev_norm:
        _, s_r, c_r = _pack_folded(handle, w, gamma, beta, bias)
        L.check(handle, lib.stlt_op_gemm_fused(handle, _stream(), L.GEMM_EPI_NORM_A, zb.data_ptr(), m,
                                               wf_rows.contiguous().data_ptr(), 2304, H, None, qkv_dev.data_ptr(), None,
                                               0, stats.data_ptr(), s_r.data_ptr(), c_r.data_ptr(), None, eps, 1))
    else:
        L.check(handle, lib.stlt_op_gemm(handle, _stream(), zb.data_ptr(), wf_rows.contiguous().data_ptr(),
                                         bias.data_ptr(), qkv_dev.data_ptr(), m, 2304, H, 1, L.GEMM_OUT_BF16, 0))
    ctx2 = torch.zeros(m, H, dtype=torch.bfloat16, device="cuda")
    L.check(handle, lib.stlt_op_attention(handle, _stream(), qkv_dev.data_ptr(), 1, mask_src.data_ptr(), n_seq, T,
                                          int(causal), ctx2.data_ptr(), 1, m))
    torch.cuda.synchronize()
    assert nerr(got, ctx2[:tokens].float()) < 1e-2


# ---- the same epilogues on the split operands of the fp32-parity mode (3 bf16 MMAs per product) ----
def _split(x: torch.Tensor) -> torch.Tensor:
    """[2, ...] bf16 planes: hi = bf16(x), lo = bf16(x - hi)."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, lo]).contiguous()


def _joined(planes: torch.Tensor) -> torch.Tensor:
    return planes[0].double() + planes[1].double()


def test_pack_folded_split_planes(handle):
    lib = L.load_library()
    g = torch.Generator(device="cuda").manual_seed(15)
    n = 3072
    w = torch.randn(n, H, device="cuda", generator=g) / math.sqrt(H)
    gamma = 1 + 0.1 * torch.randn(H, device="cuda", generator=g)
    beta = 0.1 * torch.randn(H, device="cuda", generator=g)
    bias = torch.randn(n, device="cuda", generator=g)
    wf = torch.empty(2, n, H, dtype=torch.bfloat16, device="cuda")
    s = torch.empty(n, device="cuda")
    c = torch.empty(n, device="cuda")
    L.check(handle, lib.stlt_op_pack_folded(handle, _stream(), w.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                            bias.data_ptr(), n, H, wf.data_ptr(), s.data_ptr(), c.data_ptr(), 2))
    torch.cuda.synchronize()
    assert torch.equal(wf, _split(w * gamma))
    assert nerr(_joined(wf), (w * gamma).double()) < 2e-5   # 16 mantissa bits
    assert nerr(s, _joined(wf).sum(-1)) < 1e-6
    assert nerr(c, w.double() @ beta.double() + bias.double()) < 1e-6


@pytest.mark.parametrize("m,n,gelu", [(128, 2304, 0), (384, 3072, 1), (148 * 128 + 128, 2304, 0), (640, 3072, 1)])
def test_gemm_norm_a_epilogue_split_operands(handle, m, n, gelu):
    """fp32-parity mode: act(LN(z) W^T + b) from the hi / lo split of the un-normalised stream and of the folded weights;
    the error budget of the whole forward is 1e-4, so the op has to sit near 1e-5."""
    lib = L.load_library()
    g = torch.Generator(device="cuda").manual_seed(m + n + 1)
    z = torch.randn(m, H, device="cuda", generator=g) * 1.7 + 0.3
    w = torch.randn(n, H, device="cuda", generator=g) / math.sqrt(H)
    gamma = 1 + 0.1 * torch.randn(H, device="cuda", generator=g)
    beta = 0.1 * torch.randn(H, device="cuda", generator=g)
    bias = 0.1 * torch.randn(n, device="cuda", generator=g)
    eps = 1e-5
    wf = torch.empty(2, n, H, dtype=torch.bfloat16, device="cuda")
    s = torch.empty(n, device="cuda")
    c = torch.empty(n, device="cuda")
    L.check(handle, lib.stlt_op_pack_folded(handle, _stream(), w.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                            bias.data_ptr(), n, H, wf.data_ptr(), s.data_ptr(), c.data_ptr(), 2))
    zp = _split(z)
    stats = _row_stats(z)
    out = torch.full((2, m, n), float("nan"), dtype=torch.bfloat16, device="cuda")
    L.check(handle, lib.stlt_op_gemm_fused_split(handle, _stream(), L.GEMM_EPI_NORM_A, zp.data_ptr(), m, wf.data_ptr(), n, H,
                                                 None, out.data_ptr(), None, gelu, stats.data_ptr(), s.data_ptr(),
                                                 c.data_ptr(), None, eps, 1))
    torch.cuda.synchronize()
    exact = _layer_norm64(z, gamma, beta, eps) @ w.double().T + bias.double()
    if gelu:
        exact = torch.nn.functional.gelu(exact)
    assert torch.isfinite(out.float()).all()
    assert nerr(_joined(out), exact) < 3e-5
    assert torch.equal(out[1], (_joined(out).float() - out[0].float()).to(torch.bfloat16))  # lo is the remainder of hi


@pytest.mark.parametrize("m,k,prev_norm", [(128, 768, 0), (384, 768, 1), (148 * 128 + 128, 768, 1), (256, 3072, 1),
                                           (640, 3072, 0)])
def test_gemm_resid_epilogue_split_operands(handle, m, k, prev_norm):
    """fp32-parity mode: z <- (LN(z) | z) + A W^T + b in place on split operands; the hi / lo split of the new z (the next
    GEMM's operand) and the partial row statistics are by-products."""
    lib = L.load_library()
    g = torch.Generator(device="cuda").manual_seed(m + k + prev_norm + 7)
    z = torch.randn(m, H, device="cuda", generator=g) * 1.3 - 0.2
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(H, k, device="cuda", generator=g) / math.sqrt(k)
    bias = 0.1 * torch.randn(H, device="cuda", generator=g)
    gamma = 1 + 0.1 * torch.randn(H, device="cuda", generator=g)
    beta = 0.1 * torch.randn(H, device="cuda", generator=g)
    eps = 1e-5
    ap, wp = _split(a), _split(w)
    stats_in = _row_stats(z)
    z_io = z.clone()
    zb = torch.full((2, m, H), float("nan"), dtype=torch.bfloat16, device="cuda")
    stats_out = torch.full((m, 6, 2), float("nan"), device="cuda")
    L.check(handle, lib.stlt_op_gemm_fused_split(handle, _stream(), L.GEMM_EPI_RESID, ap.data_ptr(), m, wp.data_ptr(), H, k,
                                                 bias.data_ptr(), z_io.data_ptr(), zb.data_ptr(), 0, stats_in.data_ptr(),
                                                 gamma.data_ptr(), beta.data_ptr(), stats_out.data_ptr(), eps, prev_norm))
    torch.cuda.synchronize()
    x = _layer_norm64(z, gamma, beta, eps) if prev_norm else z.double()
    ref = x + a.double() @ w.double().T + bias.double()
    assert nerr(z_io, ref) < 2e-5
    assert torch.equal(zb, _split(z_io))
    so = stats_out.double().sum(1)
    assert nerr(so[:, 0], ref.sum(-1)) < 2e-5
    assert nerr(so[:, 1], (ref * ref).sum(-1)) < 2e-5
