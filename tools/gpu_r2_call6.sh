mkdir -p gpurun_out/r2
nvidia-smi -L | head -8
timeout -s KILL 900 python -m pytest tests/test_gpu_train.py -x -q -k "two_gpu" 2>&1 | tail -5 > gpurun_out/r2/dp_test.log; cat gpurun_out/r2/dp_test.log
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29551 tests/dp_worker.py > gpurun_out/r2/dp_worker.log 2>&1; grep -E "DP_VS|DP_OK|Error|assert" gpurun_out/r2/dp_worker.log | head
timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2/bench_8gpu_v3.json 2> gpurun_out/r2/bench_8gpu_v3.err
tail -2 gpurun_out/r2/bench_8gpu_v3.err
timeout -s KILL 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2/bench_1gpu_on8box_v3.json 2>/dev/null
python - <<'PY'
import json
for n in ("bench_8gpu_v3","bench_1gpu_on8box_v3"):
    try:
        d=json.loads(open(f"gpurun_out/r2/{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", d["ms_per_step"], "clocks", d["clocks"]["sm_mhz"])
        print("  per_rank", [(r["rank"], round(r["ms_per_step"],2), round(r["kernel_ms_per_step"],2)) for r in d["per_rank"]])
        for k,v in (d.get("other_configs") or {}).items():
            print("  ", k, {kk: (round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("value","ms_per_step","e2e","error")})
    except Exception as e: print(n, "ERR", e)
PY
