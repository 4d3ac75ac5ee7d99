#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mesh.py -x -q 2>&1 | tail -15 > gpurun_out/r2c_pytest_mesh.txt
cat gpurun_out/r2c_pytest_mesh.txt
for so in libvoxb200.so libvoxb200_t3216.so libvoxb200_t1632.so libvoxb200_t3232.so libvoxb200_t1616b256.so libvoxb200_t3232b256.so; do
  echo "== $so"
  SKIP_ONESHOT=1 VOXB200_SO=$so timeout 600 python scripts/exp_tiles.py 2>&1 | grep -v "^Exception\|^Traceback\|api.py\|TypeError" | tee -a gpurun_out/r2c_exp_tiles.log
done
