mkdir -p gpurun_out/r2
timeout -s KILL 600 python -m pytest tests/test_gpu_fused_ops.py -x -q 2>&1 | tail -5
timeout -s KILL 900 python -m pytest tests/test_gpu_forward.py -x -q 2>&1 | tail -15
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2/bench_v4.json 2> gpurun_out/r2/bench_v4.err
tail -3 gpurun_out/r2/bench_v4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_v4.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["ms_per_step"], d["breakdown_ms_per_step"], d["clocks"])
print("roofline", d["roofline"]["frac"], "whole", d["whole_step_frac_of_peak"])
print("parity", d["parity"])
print("ab", {k:(round(v["ms_per_step"],2)) for k,v in d["fusion_ab"].items() if isinstance(v,dict)}, "profiled", d["profiled_pass_ms_per_step"])
print("ragged", d["ragged_batch"])
PY
