#!/bin/bash
# usage: scripts/gpurun_r2w.sh N
n=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2w_bench_n$n.json 2> gpurun_out/r2w_bench_n$n.err
tail -3 gpurun_out/r2w_bench_n$n.err
python - <<P
import json
d=json.load(open('gpurun_out/r2w_bench_n$n.json'))
print('n', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'resident', d['resident']['ms_per_step'], 'prepare', d['prepare_ms'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('phases_ms'), d['e2e'].get('table_matches_reference_golden'), 'gather', d['gather']['ms'], d['parity'])
P
