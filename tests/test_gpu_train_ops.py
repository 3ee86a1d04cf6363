"""GPU parity tests of the training-step operators, through the C ABI."""
import ctypes
import math

import pytest
import torch

from stlt_b200 import lib as L
from tests.util import nerr

pytestmark = pytest.mark.gpu

GEMM_NN, GEMM_TN_RED = 1, 2


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


# data gradient dX[M, N] = dY[M, K] * W[K, N]: the four projection shapes + a multi-wave case
NN_CASES = [(256, 768, 2304, 0), (384, 768, 768, 1), (256, 3072, 768, 1), (256, 768, 3072, 0),
            (148 * 128 * 2 + 128, 768, 768, 0)]


@pytest.mark.parametrize("m,n,k,out_kind", NN_CASES)
def test_gemm_data_gradient(handle, m, n, k, out_kind):
    lib = L.load_library()
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a = torch.randn(m, k, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(k, n, device="cuda", generator=g) / math.sqrt(k)).to(torch.bfloat16)
    out = torch.full((m, n), float("nan"), device="cuda",
                     dtype=torch.float32 if out_kind == 0 else torch.bfloat16)
    L.check(handle, lib.stlt_op_gemm_grad(handle, _stream(), GEMM_NN, a.data_ptr(), w.data_ptr(),
                                          out.data_ptr(), m, n, k, out_kind))
    ref = a.double() @ w.double()
    assert nerr(out, ref) < (5e-6 if out_kind == 0 else 6e-3)


# weight gradient dW[M, N] += dY[K, M]^T * X[K, N], K = token count (ragged allowed)
TN_CASES = [(2304, 768, 640), (768, 768, 1000), (3072, 768, 4096 + 17), (768, 3072, 85 * 64),
            (768, 768, 1), (256, 256, 148 * 64 * 3 + 5)]


@pytest.mark.parametrize("m,n,k", TN_CASES)
def test_gemm_weight_gradient(handle, m, n, k):
    lib = L.load_library()
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a = torch.randn(k, m, device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn(k, n, device="cuda", generator=g).to(torch.bfloat16)
    init = torch.randn(m, n, device="cuda", generator=g)
    out = init.clone()
    L.check(handle, lib.stlt_op_gemm_grad(handle, _stream(), GEMM_TN_RED, a.data_ptr(), b.data_ptr(),
                                          out.data_ptr(), m, n, k, 0))
    ref = init.double() + a.double().T @ b.double()
    assert nerr(out, ref) < 1e-5


@pytest.mark.parametrize("T,num_seqs,causal", [(5, 13, False), (5, 4096, False), (11, 7, False), (17, 5, True),
                                               (17, 700, True), (2, 3, True), (32, 2, True), (1, 9, False)])
@pytest.mark.parametrize("impl", [0, 1])
def test_attention_backward(handle, T, num_seqs, causal, impl):
    """dQKV of masked multi-head attention vs torch autograd on the same bf16 inputs (fp32 math)."""
    lib = L.load_library()
    g = torch.Generator(device="cuda").manual_seed(T * 1000 + num_seqs)
    tokens = num_seqs * T
    qkv = torch.randn(tokens, 2304, device="cuda", generator=g).to(torch.bfloat16)
    d_ctx = torch.randn(tokens, 768, device="cuda", generator=g).to(torch.bfloat16)
    mask_src = (torch.rand(num_seqs, T, device="cuda", generator=g) > 0.3).long()
    mask_src[:, 0] = 1  # slot / frame 0 is always valid
    d_qkv = torch.full((tokens, 2304), float("nan"), device="cuda", dtype=torch.bfloat16)
    d_bias = torch.ones(2304, device="cuda")
    L.check(handle, lib.stlt_op_attention_bwd(handle, _stream(), qkv.data_ptr(), d_ctx.data_ptr(),
                                              mask_src.data_ptr(), num_seqs, T, int(causal), d_qkv.data_ptr(), impl,
                                              d_bias.data_ptr() if impl == 1 else None))
    x = qkv.float().view(num_seqs, T, 3, 12, 64).requires_grad_(True)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))  # [N, heads, T, 64]
    scores = q @ k.transpose(-1, -2) / 8.0
    masked = (mask_src == 0).view(num_seqs, 1, 1, T).expand(num_seqs, 12, T, T)
    if causal:
        masked = masked | torch.triu(torch.ones(T, T, dtype=torch.bool, device="cuda"), diagonal=1)
    ctx = torch.softmax(scores.masked_fill(masked, float("-inf")), dim=-1) @ v
    ctx.transpose(1, 2).reshape(tokens, 768).backward(d_ctx.float())
    want = x.grad.view(tokens, 2304)
    assert torch.isfinite(d_qkv.float()).all()
    assert nerr(d_qkv, want) < 1.2e-2
    if impl == 1:  # fused in-projection bias gradient: accumulated on top of the existing value
        assert nerr(d_bias - 1.0, want.sum(0)) < 5e-3
