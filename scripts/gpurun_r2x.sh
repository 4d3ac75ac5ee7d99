#!/bin/bash
mkdir -p gpurun_out
./scripts/micro/peer_gather 180 8 2>&1 | tee gpurun_out/r2x_peer_gather.log
nvidia-smi topo -m 2>&1 | head -14 | tee -a gpurun_out/r2x_peer_gather.log
