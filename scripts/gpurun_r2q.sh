#!/bin/bash
mkdir -p gpurun_out
for so in libvoxb200.so libvoxb200_t816.so libvoxb200_t168.so libvoxb200_t88.so; do
  VOXB200_SO=$so timeout 600 python scripts/exp_slabs8.py 2>&1 | tail -9
done | tee gpurun_out/r2q_exp_slabs8.log
