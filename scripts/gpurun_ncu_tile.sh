#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:surface_tile_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r2h_tile python scripts/prof_tiles.py > gpurun_out/r2h_ncu.log 2>&1
tail -2 gpurun_out/r2h_ncu.log
