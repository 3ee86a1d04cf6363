set -x
mkdir -p gpurun_out/r2
timeout -s KILL 900 python -m pytest tests/test_gpu_fused_ops.py -x -q 2>&1 | tail -40 > gpurun_out/r2/fused_ops.log
cat gpurun_out/r2/fused_ops.log | tail -15
timeout -s KILL 900 python -m pytest tests/test_gpu_forward.py -x -q -k "fused or bf16 or batch8 or fuzz or full_size" 2>&1 | tail -30 > gpurun_out/r2/forward.log
tail -12 gpurun_out/r2/forward.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2/bench_fused_attn.json 2> gpurun_out/r2/bench_fused_attn.err
STLT_FUSED_ATTENTION=0 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2/bench_unfused_attn.json 2> gpurun_out/r2/bench_unfused_attn.err
python - <<'PY'
import json
for n in ("fused","unfused"):
    try:
        d=json.loads(open(f"gpurun_out/r2/bench_{n}_attn.json").read().strip().splitlines()[-1])
        print(n, round(d["value"]), d["ms_per_step"], d["breakdown_ms_per_step"], d["roofline"]["frac"], d["clocks"])
    except Exception as e: print(n, "ERR", e)
PY
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 200 --csv --log-file gpurun_out/r2/launches_v1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2/ncu_bench.log 2>&1
tail -3 gpurun_out/r2/ncu_bench.log
