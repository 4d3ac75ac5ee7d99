#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_large_grid.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2v_pytest_large.txt
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mesh.py tests/test_gpu_multi.py tests/test_gpu_readback.py -x -q 2>&1 | tail -5 | tee gpurun_out/r2v_pytest.txt
