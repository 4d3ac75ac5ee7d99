#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_readback.py tests/test_gpu_multi.py -x -q 2>&1 | tail -3
for w in config2 readme1024 config4; do
  timeout 600 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline > gpurun_out/r3f_bench_$w.json 2> gpurun_out/r3f_bench_$w.err
  python - <<P
import json
d=json.load(open('gpurun_out/r3f_bench_$w.json')); e=d['e2e']
print('$w', d['ms_per_step'], 'e2e', e['ms_per_step'], e.get('phases_ms'), 'dense', e['dense_readback']['ms_per_step'], 'nonzero-out', (e.get('nonzero_words_output') or {}).get('ms_per_step'))
P
done
