#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_readback.py tests/test_gpu_multi.py -x -q 2>&1 | tail -8 | tee gpurun_out/r2n_pytest_readback.txt
VOXB200_DEBUG_READBACK=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench_config4.json 2> gpurun_out/r2n_bench_config4.err
grep readback gpurun_out/r2n_bench_config4.err | tail -8; python - <<'P'
import json
d=json.load(open('gpurun_out/r2n_bench_config4.json'))
print(json.dumps(d['e2e'],indent=1))
P
