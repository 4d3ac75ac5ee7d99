#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mesh.py -x -q 2>&1 | tail -8 > gpurun_out/r2g_pytest_mesh.txt
cat gpurun_out/r2g_pytest_mesh.txt
SKIP_ONESHOT=1 timeout 600 python scripts/exp_tiles.py 2>&1 | tee gpurun_out/r2g_exp_tiles.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:surface_tile_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r2g_tile python scripts/prof_tiles.py > gpurun_out/r2g_ncu.log 2>&1
tail -2 gpurun_out/r2g_ncu.log
