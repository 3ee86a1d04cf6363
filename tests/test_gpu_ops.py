"""GPU parity tests of the single operators, called through the C ABI (stlt_op_*)."""
import ctypes
import math

import pytest
import torch

from stlt_b200 import lib as L
from tests.util import nerr

pytestmark = pytest.mark.gpu


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


def _split(x: torch.Tensor):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


def test_pack_bf16_planes(handle):
    lib = L.load_library()
    x = torch.randn(3 * 1024 + 4, device="cuda") * 3
    out = torch.empty(2 * x.numel(), dtype=torch.bfloat16, device="cuda")
    L.check(handle, lib.stlt_op_pack_bf16(handle, _stream(), x.data_ptr(), out.data_ptr(), x.numel(), 2))
    hi, lo = _split(x)
    assert torch.equal(out[: x.numel()], hi) and torch.equal(out[x.numel():], lo)


GEMM_CASES = [
    # (m, n, k, terms, out_kind, gelu) — the four projection shapes in both precision modes
    (256, 2304, 768, 1, L.GEMM_OUT_BF16, 0),
    (384, 768, 768, 1, L.GEMM_OUT_F32, 0),
    (256, 3072, 768, 1, L.GEMM_OUT_BF16, 1),
    (256, 3072, 768, 1, L.GEMM_OUT_BF16, 2),   # fast single-branch erf GELU (bf16 mode)
    (256, 768, 3072, 1, L.GEMM_OUT_F32, 0),
    (256, 2304, 768, 3, L.GEMM_OUT_F32, 0),
    (256, 3072, 768, 3, L.GEMM_OUT_BF16_SPLIT, 1),
    (128, 768, 3072, 3, L.GEMM_OUT_F32, 0),
    (148 * 128 * 2 + 128, 768, 768, 1, L.GEMM_OUT_F32, 0),   # > 2 tiles per CTA: both TMEM buffers wrap
    (148 * 128 + 256, 256, 64, 3, L.GEMM_OUT_F32, 0),         # single k-block per term
]


@pytest.mark.parametrize("m,n,k,terms,out_kind,gelu", GEMM_CASES)
def test_tcgen05_gemm(handle, m, n, k, terms, out_kind, gelu):
    lib = L.load_library()
    g = torch.Generator(device="cuda").manual_seed(m + n + k + terms)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) / math.sqrt(k)
    bias = torch.randn(n, device="cuda", generator=g)
    a_hi, a_lo = _split(a)
    w_hi, w_lo = _split(w)
    if terms == 3:
        a_p = torch.cat([a_hi, a_lo], 0).contiguous()
        w_p = torch.cat([w_hi, w_lo], 0).contiguous()
        ref = a.double() @ w.double().T + bias.double()
        tol = 3e-5
    else:
        a_p, w_p = a_hi.contiguous(), w_hi.contiguous()
        ref = a_hi.double() @ w_hi.double().T + bias.double()
        tol = 5e-6
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    if out_kind == L.GEMM_OUT_F32:
        out = torch.full((m, n), float("nan"), device="cuda")
    elif out_kind == L.GEMM_OUT_BF16:
        out = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device="cuda")
    else:
        out = torch.full((2 * m, n), float("nan"), dtype=torch.bfloat16, device="cuda")
    L.check(handle, lib.stlt_op_gemm(handle, _stream(), a_p.data_ptr(), w_p.data_ptr(), bias.data_ptr(),
                                     out.data_ptr(), m, n, k, terms, out_kind, gelu))
    torch.cuda.synchronize()
    if out_kind == L.GEMM_OUT_F32:
        got = out
    elif out_kind == L.GEMM_OUT_BF16:
        got = out.float()
        tol = 4e-3  # bf16 rounding of the output
    else:
        got = out[:m].float() + out[m:].float()
        # hi = bf16(x), lo = bf16(x - hi): lo is at most half a bf16 ulp of hi
        assert (out[m:].float().abs() <= out[:m].float().abs() * 2.0 ** -8 + 1e-30).all()
        assert nerr(out[:m].float(), ref) < 4e-3
        tol = 4e-5
    assert torch.isfinite(got).all()
    assert nerr(got, ref) < tol


def test_simt_gemm_matches_fp64(handle):
    lib = L.load_library()
    for (m, n, k, gelu) in ((37, 174, 768, 0), (130, 768, 768, 1)):
        a = torch.randn(m, k, device="cuda")
        w = torch.randn(n, k, device="cuda") / math.sqrt(k)
        b = torch.randn(n, device="cuda")
        out = torch.empty(m, n, device="cuda")
        L.check(handle, lib.stlt_op_gemm_simt(handle, _stream(), a.data_ptr(), w.data_ptr(), b.data_ptr(),
                                              out.data_ptr(), m, n, k, gelu))
        ref = a.double() @ w.double().T + b.double()
        if gelu:
            ref = torch.nn.functional.gelu(ref)
        assert nerr(out, ref) < 2e-6


def _attention_ref(qkv, masked_keys, T, causal):
    n = qkv.shape[0] // T
    q, k, v = qkv.double().view(n, T, 3, 12, 64).permute(2, 0, 3, 1, 4)  # [3][n, heads, T, d]
    s = q @ k.transpose(-1, -2) / 8.0
    mask = masked_keys.view(n, 1, 1, T).expand(n, 1, T, T).clone()
    if causal:
        mask = mask | torch.triu(torch.ones(T, T, dtype=torch.bool, device=qkv.device), 1)
    s = s.masked_fill(mask, float("-inf"))
    o = torch.softmax(s, -1) @ v
    return o.transpose(1, 2).reshape(n * T, 768)


@pytest.mark.parametrize("T,causal,n_seqs", [(5, False, 1000), (11, False, 333), (17, True, 257), (1, False, 64),
                                             (32, True, 31), (7, True, 50)])
@pytest.mark.parametrize("flavour", ["simt_f32", "mma_bf16", "mma_split"])
def test_attention(handle, T, causal, n_seqs, flavour):
    """simt_f32: fp32 QKV on CUDA cores (cross-check); mma_bf16: bf16 mode; mma_split: fp32-parity mode
    (hi/lo bf16 planes in and out, 3-term split products on the tensor cores)."""
    lib = L.load_library()
    g = torch.Generator(device="cuda").manual_seed(T * 7 + n_seqs)
    tokens = n_seqs * T
    rows = tokens + 5
    qkv = torch.randn(tokens, 2304, device="cuda", generator=g)
    if flavour == "mma_bf16":
        qkv_in = qkv.to(torch.bfloat16)
        qkv = qkv_in.float()
        planes, is_bf16, tol = 1, 1, 4e-3
    elif flavour == "mma_split":
        hi, lo = _split(qkv)
        qkv_in = torch.zeros(2 * rows, 2304, dtype=torch.bfloat16, device="cuda")
        qkv_in[:tokens] = hi
        qkv_in[rows: rows + tokens] = lo
        planes, is_bf16, tol = 2, 1, 3e-5
    else:
        qkv_in = qkv
        planes, is_bf16, tol = 2, 0, 2e-5
    mask_src = torch.randint(0, 3, (n_seqs, T), device="cuda", generator=g)
    mask_src[:, 0] = 2  # first key always valid (CLS slot / first frame)
    mask_src = mask_src.view(-1).contiguous()
    out = torch.zeros(planes * rows, 768, dtype=torch.bfloat16, device="cuda")
    L.check(handle, lib.stlt_op_attention(handle, _stream(), qkv_in.data_ptr(), is_bf16, mask_src.data_ptr(),
                                          n_seqs, T, int(causal), out.data_ptr(), planes, rows))
    torch.cuda.synchronize()
    got = out[:tokens].float()
    if planes == 2:
        got = got + out[rows: rows + tokens].float()
    ref = _attention_ref(qkv, mask_src == 0, T, causal)
    assert nerr(got, ref) < tol


@pytest.mark.parametrize("T,causal,n_seqs", [(65, True, 9), (128, False, 5), (200, True, 7), (256, True, 3),
                                             (256, False, 2), (97, False, 40)])
@pytest.mark.parametrize("flavour", ["mma_bf16", "mma_split"])
def test_attention_long_sequences(handle, T, causal, n_seqs, flavour):
    """65..256 tokens per sequence (attention_long.cu: online softmax over 64-key blocks) vs fp64 torch; random
    key masks, so whole key blocks can be masked for some rows."""
    lib = L.load_library()
    g = torch.Generator(device="cuda").manual_seed(T * 11 + n_seqs)
    tokens = n_seqs * T
    rows = tokens + 5
    qkv = torch.randn(tokens, 2304, device="cuda", generator=g)
    if flavour == "mma_bf16":
        qkv_in = qkv.to(torch.bfloat16)
        qkv = qkv_in.float()
        planes, tol = 1, 4e-3
    else:
        hi, lo = _split(qkv)
        qkv_in = torch.zeros(2 * rows, 2304, dtype=torch.bfloat16, device="cuda")
        qkv_in[:tokens] = hi
        qkv_in[rows: rows + tokens] = lo
        planes, tol = 2, 3e-5
    mask_src = torch.randint(0, 3, (n_seqs, T), device="cuda", generator=g)
    mask_src[:, 0] = 2  # first key always valid (first frame)
    mask_src[0, 1:70] = 0  # a run of masked keys longer than a key block
    mask_src = mask_src.view(-1).contiguous()
    out = torch.zeros(planes * rows, 768, dtype=torch.bfloat16, device="cuda")
    L.check(handle, lib.stlt_op_attention(handle, _stream(), qkv_in.data_ptr(), 1, mask_src.data_ptr(),
                                          n_seqs, T, int(causal), out.data_ptr(), planes, rows))
    torch.cuda.synchronize()
    got = out[:tokens].float()
    if planes == 2:
        got = got + out[rows: rows + tokens].float()
    ref = _attention_ref(qkv, mask_src == 0, T, causal)
    assert torch.isfinite(got).all()
    assert nerr(got, ref) < tol


@pytest.mark.parametrize("rows,with_y,planes", [(1000, True, 2), (77, False, 1), (3, True, 1)])
def test_add_layer_norm(handle, rows, with_y, planes):
    lib = L.load_library()
    x = torch.randn(rows, 768, device="cuda") * 2 + 0.3
    y = torch.randn(rows, 768, device="cuda") if with_y else None
    gam = torch.randn(768, device="cuda")
    bet = torch.randn(768, device="cuda")
    out = torch.empty(rows, 768, device="cuda")
    prow = rows + 3
    outb = torch.zeros(planes * prow, 768, dtype=torch.bfloat16, device="cuda")
    for eps in (1e-5, 1e-12):
        L.check(handle, lib.stlt_op_add_ln(handle, _stream(), x.data_ptr(), y.data_ptr() if with_y else None,
                                           gam.data_ptr(), bet.data_ptr(), eps, rows, out.data_ptr(),
                                           outb.data_ptr(), planes, prow))
        s = x.double() + (y.double() if with_y else 0)
        ref = torch.nn.functional.layer_norm(s, (768,), gam.double(), bet.double(), eps)
        assert nerr(out, ref) < 2e-6
        assert torch.equal(outb[:rows], out.to(torch.bfloat16))
        if planes == 2:
            assert torch.equal(outb[prow: prow + rows], (out - out.to(torch.bfloat16).float()).to(torch.bfloat16))
