#!/bin/bash
mkdir -p gpurun_out
for e in 0 2 3 4 8; do VOXB200_PREZERO_EARLY=$e NDEV=1 timeout 120 python scripts/exp_multi_e2e.py 2>&1 | grep "^\["; done | tee gpurun_out/r2u_exp_prezero_early.log
