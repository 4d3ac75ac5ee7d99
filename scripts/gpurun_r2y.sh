#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_readback.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2y_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2y_bench_config4.json 2> gpurun_out/r2y_bench_config4.err
tail -2 gpurun_out/r2y_bench_config4.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2y_bench_config4.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['phases_ms'], d['e2e']['nonzero_words_output'], d['e2e']['dense_readback'], d['e2e']['host_table_matches_device_table'])
P
