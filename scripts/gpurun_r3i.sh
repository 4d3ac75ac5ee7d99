#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in config2 config4 readme1024; do
timeout 600 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline > gpurun_out/r3i_bench_$w.json 2> gpurun_out/r3i_bench_$w.err
python -c "
import json;d=json.load(open('gpurun_out/r3i_bench_$w.json'));print('$w',d['value'],d['ms_per_step'],'one_shot',d['one_shot']['ms_per_step'],d['roofline']['phases_ms'],d['parity'])"
done
