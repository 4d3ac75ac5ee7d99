#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mesh.py -x -q 2>&1 | tail -5 > gpurun_out/r2d_pytest_mesh.txt
cat gpurun_out/r2d_pytest_mesh.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:surface_tile_red --launch-skip 3 -c 1 -f -o gpurun_out/r2d_tile_red python scripts/prof_tiles.py > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log
