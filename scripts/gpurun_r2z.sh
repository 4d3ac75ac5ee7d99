#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2z_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2z_bench_config4.json 2> gpurun_out/r2z_bench_config4.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2z_bench_config4.json'))
print(d['value'], d['ms_per_step'], 'resident', d['resident']['ms_per_step'], 'one_shot', d['one_shot']['ms_per_step'], d['roofline']['phases_ms'], d['parity'])
P
