#!/bin/bash
# 2 GPUs: multi-GPU C-ABI tests + bench at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15 > gpurun_out/r2i_pytest_multi.txt
cat gpurun_out/r2i_pytest_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err
tail -5 gpurun_out/r2i_bench_n2.err; cat gpurun_out/r2i_bench_n2.json
