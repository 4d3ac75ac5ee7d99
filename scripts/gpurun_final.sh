#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/final_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; python -c "
import json;d=json.load(open('gpurun_out/final_bench.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['e2e']['ms_per_step'],d['gpu_launches'],d['clocks'],d['parity'])"
