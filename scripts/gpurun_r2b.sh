#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mesh.py -x -q 2>&1 | tail -15 > gpurun_out/r2b_pytest_mesh.txt
cat gpurun_out/r2b_pytest_mesh.txt
timeout 600 python scripts/exp_tiles.py > gpurun_out/r2b_exp_tiles_red.log 2>&1
cat gpurun_out/r2b_exp_tiles_red.log
SKIP_ONESHOT=1 VOXB200_SO=libvoxb200_bytes.so timeout 600 python scripts/exp_tiles.py > gpurun_out/r2b_exp_tiles_bytes.log 2>&1
cat gpurun_out/r2b_exp_tiles_bytes.log
SKIP_ONESHOT=1 timeout 600 python scripts/exp_tiles.py icosphere:224:512 1024 > gpurun_out/r2b_exp_tiles_c3mesh.log 2>&1
cat gpurun_out/r2b_exp_tiles_c3mesh.log
timeout 600 python scripts/exp_tiles.py bunny 1024 > gpurun_out/r2b_exp_tiles_bunny.log 2>&1
cat gpurun_out/r2b_exp_tiles_bunny.log
