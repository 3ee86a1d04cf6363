"""Timing decomposition of the in-projection + attention kernel (csrc/gemm_qkv_attn.cu) at the bench shapes, next to
the unfused pair it replaces (in-projection GEMM + attention kernel). STLT_QKV_ATTN_DEBUG variants switch parts of
the epilogue off (results are garbage then): 1 = no attention math, 2 = no Q/K/V tile stores, 4 = no statistics loads, 8 = a 2 us sleep per unit in the epilogue.

    python tools/bench_qkv_attention.py [--batch 4096]
"""
import argparse
import ctypes
import math
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stlt_b200 import lib as L  # noqa: E402

H = 768


def make_handle():
    lib = L.load_library()
    dims = L.StltDims(768, 12, 0, 0, 4, 174, 256, 5, 1e-12, 1e-5)
    h = ctypes.c_void_p()
    L.check(None, lib.stlt_create(ctypes.byref(dims), ctypes.byref(h)))
    return h


def time_it(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    args = ap.parse_args()
    lib = L.load_library()
    stream = torch.cuda.current_stream().cuda_stream
    for name, T, n_seq, causal in (("spatial", 5, args.batch * 17, 0), ("temporal", 17, args.batch, 1)):
        tokens = T * n_seq
        m = (tokens + 127) // 128 * 128
        z = torch.randn(m, H, device="cuda")
        zb = z.to(torch.bfloat16)
        w = torch.randn(2304, H, device="cuda") / math.sqrt(H)
        bias = torch.randn(2304, device="cuda") * 0.1
        gamma = torch.ones(H, device="cuda")
        beta = torch.zeros(H, device="cuda")
        stats = torch.stack([z.view(m, 6, 128).sum(-1), (z * z).view(m, 6, 128).sum(-1)], -1).contiguous()
        mask = torch.ones(tokens, dtype=torch.int64, device="cuda")
        wh = torch.empty(2304, H, dtype=torch.bfloat16, device="cuda")
        sh, ch = torch.empty(2304, device="cuda"), torch.empty(2304, device="cuda")
        h0 = make_handle()
        L.check(h0, lib.stlt_op_pack_folded(h0, stream, w.data_ptr(), gamma.data_ptr(), beta.data_ptr(), bias.data_ptr(),
                                            2304, H, wh.data_ptr(), sh.data_ptr(), ch.data_ptr(), 1))
        ctx = torch.empty(m, H, dtype=torch.bfloat16, device="cuda")
        qkv = torch.empty(m, 2304, dtype=torch.bfloat16, device="cuda")
        wp = w.to(torch.bfloat16).contiguous()
        flops = 2.0 * m * 2304 * H
        ms_g = time_it(lambda: L.check(h0, lib.stlt_op_gemm(h0, stream, zb.data_ptr(), wp.data_ptr(), bias.data_ptr(),
                                                            qkv.data_ptr(), m, 2304, H, 1, L.GEMM_OUT_BF16, 0)))
        ms_a = time_it(lambda: L.check(h0, lib.stlt_op_attention(h0, stream, qkv.data_ptr(), 1, mask.data_ptr(), n_seq, T,
                                                                 causal, ctx.data_ptr(), 1, m)))
        print(f"{name}: T={T} tokens={tokens}  unfused: in-projection {ms_g:.3f} ms ({flops / ms_g / 1e9:.0f} TFLOP/s) + "
              f"attention {ms_a:.3f} ms = {ms_g + ms_a:.3f} ms")
        for dbg in (0, 1, 9, 7):
            os.environ["STLT_QKV_ATTN_DEBUG"] = str(dbg)
            h = make_handle()
            for prev in (1,):
                ms = time_it(lambda: L.check(h, lib.stlt_op_qkv_attention(
                    h, stream, zb.data_ptr(), m, tokens, wh.data_ptr(), sh.data_ptr(), ch.data_ptr(),
                    stats.data_ptr() if prev else None, 1e-5, mask.data_ptr(), n_seq, T, causal, ctx.data_ptr())))
                print(f"  fused debug={dbg} prev_norm={prev}: {ms:.3f} ms ({flops / ms / 1e9:.0f} TFLOP/s of in-projection FLOPs)")
            lib.stlt_destroy(h)
        os.environ["STLT_QKV_ATTN_DEBUG"] = "0"
        lib.stlt_destroy(h0)


if __name__ == "__main__":
    main()
