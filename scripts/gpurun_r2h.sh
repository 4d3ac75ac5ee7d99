#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench_config4.json 2> gpurun_out/r2h_bench_config4.err
tail -3 gpurun_out/r2h_bench_config4.err; cat gpurun_out/r2h_bench_config4.json
timeout 600 python bench.py --steps 20 --warmup 5 --workload config2 --no-cpu-baseline > gpurun_out/r2h_bench_config2.json 2> gpurun_out/r2h_bench_config2.err
tail -3 gpurun_out/r2h_bench_config2.err; cat gpurun_out/r2h_bench_config2.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2h_pytest_gpu.txt
cat gpurun_out/r2h_pytest_gpu.txt
