mkdir -p gpurun_out/r2
timeout -s KILL 900 python -m pytest tests/test_cacnf.py -x -q 2>&1 | tail -8
timeout -s KILL 900 python bench.py --workload cacnf --batch 2048 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2/bench_cacnf_v5.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_cacnf_v5.json").read().strip().splitlines()[-1])
print("cacnf", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["ms_per_step"], d["breakdown_ms_per_step"], d["roofline"]["frac"], d["gpu_launches"])
PY
