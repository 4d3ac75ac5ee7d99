#!/bin/bash
# usage (on the GPU box, via gpurun): scripts/gpurun_round2.sh <tag>  — round 2: tests, bench lines, ncu launch lists + full captures, sanitizer
tag=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${tag}_pytest_gpu.txt
cat gpurun_out/${tag}_pytest_gpu.txt
for w in config4 config2 config3 readme1024 readme2048; do
  timeout 600 python bench.py --steps 20 --warmup 5 --workload $w > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
for w in config4 config2 config3; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_$w.csv python bench.py --steps 3 --warmup 3 --workload $w --no-cpu-baseline > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'surface_tile_kernel|tile_scatter_kernel|tile_count_kernel|tile_plan_kernel' --launch-skip 4 -c 4 -f -o gpurun_out/${tag}_config4_full python bench.py --steps 3 --warmup 3 --workload config4 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'solid_tri_kernel|solid_fill_kernel' --launch-skip 8 -c 2 -f -o gpurun_out/${tag}_config3_full python bench.py --steps 3 --warmup 3 --workload config3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'nz_count_kernel|nz_write_kernel' -c 2 -f -o gpurun_out/${tag}_readback_full python scripts/exp_readback.py > /dev/null 2>&1
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize.py > gpurun_out/${tag}_sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize.py > gpurun_out/${tag}_sanitizer_racecheck.log 2>&1
tail -3 gpurun_out/${tag}_sanitizer_memcheck.log; tail -3 gpurun_out/${tag}_sanitizer_racecheck.log
for w in config4 config2 config3 readme1024 readme2048; do python -c "
import json;d=json.load(open('gpurun_out/${tag}_bench_$w.json'));print('$w',d['value'],d['ms_per_step'],d['roofline']['phases_ms'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'cpu',d['cpu_baseline'] and d['cpu_baseline']['value'],'refgpu',d['ref_gpu_baseline'] and d['ref_gpu_baseline'].get('ms_per_step'))"; done
cat gpurun_out/${tag}_bench_reference.json
