mkdir -p gpurun_out/r2
for mt in 32 8 32 8; do
STLT_FUSED_ATTENTION_MAX_T=$mt timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-extras --no-parity > gpurun_out/r2/bench_maxt_$mt.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r2/bench_maxt_$mt.json").read().strip().splitlines()[-1])
print("max_t=$mt", round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["breakdown_ms_per_step"], d["clocks"]["sm_mhz"], {k:round(v["ms_per_step"],2) for k,v in d["fusion_ab"].items() if isinstance(v,dict)})
PY
done
