#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python scripts/exp_multi_e2e.py
  VOXB200_NO_PREZERO=1 timeout 300 python scripts/exp_multi_e2e.py
  VOXB200_HOST_THREADS=4 timeout 300 python scripts/exp_multi_e2e.py
  VOXB200_HOST_THREADS=16 timeout 300 python scripts/exp_multi_e2e.py
  VOXB200_READBACK=dense timeout 300 python scripts/exp_multi_e2e.py ) 2>&1 | grep "^\[" | tee gpurun_out/r2r_exp_multi_e2e.log
