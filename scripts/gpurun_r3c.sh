#!/bin/bash
mkdir -p gpurun_out
for so in libvoxb200.so libvoxb200_pers.so; do
  SKIP_ONESHOT=1 VOXB200_SO=$so timeout 300 python scripts/exp_tiles.py 2>&1 | grep -E "prepared mesh|identical" | sed "s/^/$so: /"
  SKIP_ONESHOT=1 VOXB200_SO=$so timeout 300 python scripts/exp_tiles.py bunny 1024 2>&1 | grep -E "prepared mesh" | sed "s/^/$so bunny1024: /"
  SKIP_ONESHOT=1 VOXB200_SO=$so timeout 300 python scripts/exp_tiles.py icosphere:224:512 1024 2>&1 | grep -E "prepared mesh" | sed "s/^/$so ico224@1024: /"
done | tee gpurun_out/r3c_exp_persistent.log
VOXB200_SO=libvoxb200_pers.so PARTS=8 timeout 300 python scripts/exp_slabs8.py 2>&1 | tail -3 | tee -a gpurun_out/r3c_exp_persistent.log
