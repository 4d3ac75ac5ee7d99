#!/bin/bash
mkdir -p gpurun_out
( NDEV=1,2 timeout 120 python scripts/exp_multi_e2e.py
  NDEV=2,1,2 VOXB200_NO_PREZERO=1 timeout 120 python scripts/exp_multi_e2e.py ) 2>&1 | grep -E "^\[|Error|error" | tee gpurun_out/r2s_exp_multi_e2e.log
timeout 600 python -m pytest tests/test_gpu_readback.py tests/test_gpu_multi.py -x -q 2>&1 | tail -5 | tee gpurun_out/r2s_pytest.txt
