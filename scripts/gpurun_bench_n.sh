#!/bin/bash
# usage (via gpurun --gpus N): scripts/gpurun_bench_n.sh N [tag] — the bench line at N GPUs under torchrun
n=$1; tag=${2:-r2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
tail -3 gpurun_out/${tag}_bench_n$n.err
python - <<P
import json
d=json.load(open('gpurun_out/${tag}_bench_n$n.json')); e=d['e2e']
print('n', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'resident', d['resident']['ms_per_step'], 'prepare', d['prepare_ms'], 'e2e', e['ms_per_step'], e.get('phases_ms'), e.get('table_matches_reference_golden'), 'd2h', e['d2h_bytes_per_step'], e.get('readback'), 'gather', d['gather']['ms'], d['parity'])
P
