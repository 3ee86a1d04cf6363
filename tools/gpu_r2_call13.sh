mkdir -p gpurun_out/r2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 > gpurun_out/r2/smoke.log; cat gpurun_out/r2/smoke.log
timeout -s KILL 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2/pytest_gpu_all.log; tail -3 gpurun_out/r2/pytest_gpu_all.log
timeout -s KILL 600 compute-sanitizer --tool memcheck python tests/sanitizer_smoke.py > gpurun_out/r2/memcheck.log 2>&1; tail -2 gpurun_out/r2/memcheck.log
timeout -s KILL 900 python tools/measure_traffic.py > gpurun_out/r2/traffic.log 2>&1; cp profiles/roofline_traffic.json gpurun_out/r2/roofline_traffic.json; grep -E '"bf16"|"fp32"|digest' gpurun_out/r2/roofline_traffic.json
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second --clock-control none -c 300 --csv --log-file gpurun_out/r2/launches_v6.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-secondary --no-parity --no-graphs > gpurun_out/r2/ncu_list.log 2>&1; tail -1 gpurun_out/r2/ncu_list.log | cut -c1-200
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k 'regex:gemm_tcgen05|qkv_attention' -s 4 -c 4 -o gpurun_out/r2/layer_v6 python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-secondary --no-parity --no-graphs > gpurun_out/r2/ncu_full2.log 2>&1; tail -1 gpurun_out/r2/ncu_full2.log
timeout -s KILL 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_v6.json 2> gpurun_out/r2/bench_v6.err; tail -c 1500 gpurun_out/r2/bench_v6.json
timeout -s KILL 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2/bench_ref_v6.json 2>/dev/null; tail -c 800 gpurun_out/r2/bench_ref_v6.json
