mkdir -p gpurun_out/r2
timeout -s KILL 600 python -m pytest tests/test_gpu_fused_ops.py -x -q 2>&1 | tail -5
timeout -s KILL 900 python -m pytest tests/test_gpu_forward.py tests/test_cacnf.py -x -q 2>&1 | tail -6
timeout -s KILL 600 python tools/bench_qkv_attention.py 2>&1 | tee gpurun_out/r2/qkv_attn_decomp3.txt
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-secondary > gpurun_out/r2/bench_v5.json 2> gpurun_out/r2/bench_v5.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_v5.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["ms_per_step"], d["breakdown_ms_per_step"], d["clocks"])
print("roofline", d["roofline"]["frac"], "whole", d["whole_step_frac_of_peak"], "parity", d["parity"]["bf16"], d["parity"]["bf16_top1_agree"], d["parity"]["ok"])
print("ab", {k:(round(v["ms_per_step"],2)) for k,v in d["fusion_ab"].items() if isinstance(v,dict)})
print("ragged", d["ragged_batch"]["ms_per_step"], d["ragged_batch"]["padded_grid_ms_per_step"])
PY
