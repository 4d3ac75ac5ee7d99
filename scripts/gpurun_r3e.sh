#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_multi.py tests/test_gpu_large_grid.py -x -q 2>&1 | tail -4
SKIP_ONESHOT=1 timeout 300 python scripts/exp_tiles.py 2>&1 | grep -E "prepared mesh|update"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3e_launches.csv -k regex:'tile_plan|tile_count|tile_scatter' -c 12 python scripts/prof_tiles.py > /dev/null 2>&1
python scripts/ncu_summary.py launches gpurun_out/r3e_launches.csv
