#!/bin/bash
mkdir -p gpurun_out
for so in libvoxb200.so libvoxb200_mb5.so libvoxb200_mb4.so libvoxb200_mb3.so libvoxb200_mb2.so; do
  SKIP_ONESHOT=1 VOXB200_SO=$so timeout 300 python scripts/exp_tiles.py 2>&1 | grep -E "prepared mesh" | sed "s/^/$so: /"
done | tee -a gpurun_out/r3a_exp_occupancy.log
