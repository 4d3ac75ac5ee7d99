#!/bin/bash
mkdir -p gpurun_out
for s in 8 16 32 128 256; do echo "slices $s"; VOXB200_READBACK_SLICES=$s timeout 900 python scripts/exp_readback.py 2>&1 | grep -E "threads=(4|8|12)"; done | tee gpurun_out/r2m_exp_readback_slices.log
