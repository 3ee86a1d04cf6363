#!/usr/bin/env python
"""Micro-benchmark of the attention kernels through stlt_op_attention (T tokens per sequence, bf16)."""
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from stlt_b200 import lib as L  # noqa: E402

lib = L.load_library()
dims = L.StltDims(768, 12, 0, 0, 4, 174, 256, 5, 1e-12, 1e-5)
h = ctypes.c_void_p()
L.check(None, lib.stlt_create(ctypes.byref(dims), ctypes.byref(h)))
s = torch.cuda.current_stream().cuda_stream
for T, num_seqs, causal in ((5, 69632, 0), (17, 4096, 1), (11, 69632, 0), (33, 2048, 0)):
    tokens = T * num_seqs
    qkv = torch.randn(tokens, 2304, device="cuda").to(torch.bfloat16)
    mask = torch.ones(tokens, dtype=torch.int64, device="cuda")
    out = torch.empty(tokens, 768, dtype=torch.bfloat16, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def run():
        L.check(h, lib.stlt_op_attention(h, s, qkv.data_ptr(), 1, mask.data_ptr(), num_seqs, T, causal, out.data_ptr(), 1,
                                         tokens))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    times = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    ms = sorted(times)[len(times) // 2]
    gb = tokens * (4608 + 1536) / 1e9
    print(f"T={T:2d} seqs={num_seqs:6d}: {ms:.4f} ms, {gb / ms * 1e3:.0f} GB/s algorithmic")
