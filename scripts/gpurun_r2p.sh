#!/bin/bash
# 8 GPUs: multi-GPU C-ABI tests + bench at N=8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5 | tee gpurun_out/r2p_pytest_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2p_bench_n8.json 2> gpurun_out/r2p_bench_n8.err
tail -5 gpurun_out/r2p_bench_n8.err; cat gpurun_out/r2p_bench_n8.json
timeout 600 ./scripts/micro/host_bw 1024 8 2>&1 | grep -E "D2H|H2D" | tee gpurun_out/r2p_host_bw_n8.log
