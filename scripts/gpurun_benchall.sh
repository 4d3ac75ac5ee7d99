#!/bin/bash
# usage: scripts/gpurun_benchall.sh <tag> — the bench lines of every workload + the reference arm
tag=${1:-r2}
mkdir -p gpurun_out
for w in config4 config2 config3 readme1024 readme2048; do
  timeout 600 python bench.py --steps 20 --warmup 5 --workload $w > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  python -c "
import json;d=json.load(open('gpurun_out/${tag}_bench_$w.json'));e=d['e2e'];print('$w',d['value'],d['ms_per_step'],d['roofline']['phases_ms'],'frac',d['roofline']['frac'],'step',d['roofline']['step'],'e2e',e['value'],e['ms_per_step'],e.get('phases_ms'),'nonzero-out',(e.get('nonzero_words_output') or {}).get('ms_per_step'),'dense',e['dense_readback']['ms_per_step'],'cpu',d['cpu_baseline'] and d['cpu_baseline']['value'],'refgpu',d['ref_gpu_baseline'] and d['ref_gpu_baseline'].get('ms_per_step'),'launches',d['gpu_launches'])"
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
cut -c1-300 gpurun_out/${tag}_bench_reference.json
