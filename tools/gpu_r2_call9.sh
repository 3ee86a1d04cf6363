mkdir -p gpurun_out/r2
timeout -s KILL 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2/pytest_gpu_all.log; cat gpurun_out/r2/pytest_gpu_all.log
