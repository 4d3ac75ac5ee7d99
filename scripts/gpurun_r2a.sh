#!/bin/bash
# round 2, first GPU call: the new prepared-mesh tests, the tile experiment, the reference GPU kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mesh.py -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest_mesh.txt
cat gpurun_out/r2a_pytest_mesh.txt
timeout 600 python scripts/exp_tiles.py > gpurun_out/r2a_exp_tiles.log 2>&1
cat gpurun_out/r2a_exp_tiles.log
timeout 600 python scripts/exp_refgpu.py > gpurun_out/r2a_refgpu.log 2>&1
cat gpurun_out/r2a_refgpu.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "box or golden" 2>&1 | tail -5 > gpurun_out/r2a_pytest_parity.txt
cat gpurun_out/r2a_pytest_parity.txt
