for v in 0 1 0 1; do
if [ $v = 1 ]; then export STLT_RESID_STAGED=1; else unset STLT_RESID_STAGED; fi
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-extras --no-parity > /tmp/b.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("/tmp/b.json").read().strip().splitlines()[-1])
print("staged=$v", round(d["value"]), d["ms_per_step"], round(d["fusion_ab"]["default"]["ms_per_step"],2), d["clocks"]["sm_mhz"])
PY
done
