mkdir -p gpurun_out/r2
for aw in 0 48 0 48; do
STLT_DEBUG_A_WRAP=$aw timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-extras --no-parity > gpurun_out/r2/bench_aw_$aw.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r2/bench_aw_$aw.json").read().strip().splitlines()[-1])
print("a_wrap=$aw", round(d["value"]), d["ms_per_step"], d["breakdown_ms_per_step"], d["clocks"]["sm_mhz"], "default_ab", round(d["fusion_ab"]["default"]["ms_per_step"],2))
PY
done
