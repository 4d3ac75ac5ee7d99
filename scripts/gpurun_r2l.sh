#!/bin/bash
# sparse read-back + slimmer solid kernels: tests, then bench of configs 4 and 3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_readback.py tests/test_gpu_multi.py -x -q 2>&1 | tail -8 | tee gpurun_out/r2l_pytest_readback.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "solid or golden" 2>&1 | tail -8 | tee gpurun_out/r2l_pytest_solid.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_bench_config4.json 2> gpurun_out/r2l_bench_config4.err
tail -3 gpurun_out/r2l_bench_config4.err; cat gpurun_out/r2l_bench_config4.json
timeout 600 python bench.py --steps 20 --warmup 5 --workload config3 --no-cpu-baseline > gpurun_out/r2l_bench_config3.json 2> gpurun_out/r2l_bench_config3.err
tail -3 gpurun_out/r2l_bench_config3.err; cat gpurun_out/r2l_bench_config3.json
