timeout -s KILL 600 python tools/bench_qkv_attention.py 2>&1 | tee gpurun_out/r2/qkv_attn_decomp2.txt
