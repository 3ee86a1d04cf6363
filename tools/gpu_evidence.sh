#!/bin/bash
# Reproduces the round-2 evidence under profiles/ on a B200 box (what was run through `gpurun -- bash tools/gpu_evidence.sh`;
# about 6 GPU-minutes). Outputs land in gpurun_out/r2/; the summaries kept under profiles/ were written from them with
# tools/ncu_summary.py. Every step is wrapped in `timeout -s KILL` so that a hung kernel cannot hold the box.
mkdir -p gpurun_out/r2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 > gpurun_out/r2/smoke.log; cat gpurun_out/r2/smoke.log
timeout -s KILL 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2/pytest_gpu_all.log; tail -3 gpurun_out/r2/pytest_gpu_all.log
timeout -s KILL 600 compute-sanitizer --tool memcheck python tests/sanitizer_smoke.py > gpurun_out/r2/memcheck.log 2>&1; tail -2 gpurun_out/r2/memcheck.log
timeout -s KILL 600 compute-sanitizer --tool racecheck python tests/sanitizer_smoke.py > gpurun_out/r2/racecheck.log 2>&1; grep -c hazard gpurun_out/r2/racecheck.log
# roofline.traffic: DRAM bytes per GEMM-class launch, stored with the digest of the kernel sources (bench.py refuses a stale one)
timeout -s KILL 900 python tools/measure_traffic.py > gpurun_out/r2/traffic.log 2>&1; cp profiles/roofline_traffic.json gpurun_out/r2/roofline_traffic.json
# launch list of the bench step (shares per kernel) and a full capture of the four GEMM-class kernels of one spatial layer
NCU_BENCH="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-secondary --no-parity --no-graphs"
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second \
  --clock-control none -c 300 --csv --log-file gpurun_out/r2/launches.csv $NCU_BENCH > gpurun_out/r2/ncu_list.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tcgen05|qkv_attention' -s 4 -c 4 \
  -o gpurun_out/r2/layer $NCU_BENCH > gpurun_out/r2/ncu_layer.log 2>&1
# the HBM-bound stage kernels (embedding with the pad-skipping scatter, frame embedding, gathers, planning kernels)
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k 'regex:embed_kernel|gather_frames|plan_|gather_rows' -s 9 -c 9 \
  -o gpurun_out/r2/stages $NCU_BENCH > gpurun_out/r2/ncu_stages.log 2>&1
# timing decomposition of the attention-fused in-projection
timeout -s KILL 600 python tools/bench_qkv_attention.py > gpurun_out/r2/qkv_attn_decomp.txt 2>&1
# the two arms exactly as the driver runs them
timeout -s KILL 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_default.json 2> gpurun_out/r2/bench_default.err
timeout -s KILL 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2/bench_reference.json 2>/dev/null
tail -c 400 gpurun_out/r2/bench_default.json
