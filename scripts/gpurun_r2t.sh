#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_readback.py tests/test_gpu_binvox.py tests/test_gpu_multi.py -x -q 2>&1 | tail -5 | tee gpurun_out/r2t_pytest.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "extract or cli or upload" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_bench_config4.json 2> gpurun_out/r2t_bench_config4.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2t_bench_config4.json'))
print(d['value'], d['ms_per_step'], 'prepare', d['prepare_ms'], 'resident', d['resident']['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['phases_ms'], d['parity'])
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t_launches.csv -k regex:'tile_plan|extract_scan|nz_' -c 30 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2t_launches.csv
