#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mesh.py -x -q -k "solid or golden or fuzz or degenerate or tiny or odd" 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --workload config3 --no-cpu-baseline > gpurun_out/r3h_bench_config3.json 2> gpurun_out/r3h_bench_config3.err
python -c "
import json;d=json.load(open('gpurun_out/r3h_bench_config3.json'));print(d['value'],d['ms_per_step'],d['roofline']['phases_ms'],d['roofline']['step'],d['parity'])"
