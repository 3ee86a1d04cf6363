mkdir -p gpurun_out/r2
timeout -s KILL 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_data.py -x -q 2>&1 | tail -3
timeout -s KILL 900 python tools/measure_traffic.py > gpurun_out/r2/traffic.log 2>&1; cp profiles/roofline_traffic.json gpurun_out/r2/roofline_traffic.json; grep -E '"bf16"|"fp32"|digest' gpurun_out/r2/roofline_traffic.json
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second --clock-control none -c 300 --csv --log-file gpurun_out/r2/launches_v7.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-secondary --no-parity --no-graphs > gpurun_out/r2/ncu_list.log 2>&1
timeout -s KILL 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2/bench_v7.json 2> gpurun_out/r2/bench_v7.err; tail -c 300 gpurun_out/r2/bench_v7.json
