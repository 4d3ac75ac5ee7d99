#!/bin/bash
# new solid kernels: parity tests + ncu of the two kernels on config 3
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "solid or golden" 2>&1 | tail -8 | tee gpurun_out/r2k_pytest_solid.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'solid_tri_kernel|solid_fill_kernel' --launch-skip 8 -c 2 -f -o gpurun_out/r2k_config3_full python bench.py --steps 3 --warmup 3 --workload config3 --no-cpu-baseline > gpurun_out/r2k_ncu.log 2>&1
tail -3 gpurun_out/r2k_ncu.log
