set -x
mkdir -p gpurun_out/r2
timeout -s KILL 600 python -m pytest tests/test_gpu_fused_ops.py -x -q 2>&1 | tail -15
timeout -s KILL 900 python -m pytest tests/test_gpu_forward.py -x -q 2>&1 | tail -15
timeout -s KILL 600 python tools/bench_qkv_attention.py > gpurun_out/r2/qkv_attn_decomp.txt 2>&1
cat gpurun_out/r2/qkv_attn_decomp.txt
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2/bench_v2.json 2> gpurun_out/r2/bench_v2.err
tail -c 6000 gpurun_out/r2/bench_v2.json; tail -5 gpurun_out/r2/bench_v2.err
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:qkv_attention -s 3 -c 1 -o gpurun_out/r2/qkv_attn_v2 python tools/bench_qkv_attention.py > gpurun_out/r2/ncu_full.log 2>&1
tail -3 gpurun_out/r2/ncu_full.log; ls -la gpurun_out/r2/
