#!/bin/bash
# usage (on the GPU box, via gpurun): scripts/gpurun_round.sh <tag>  — tests, bench lines, ncu launch lists + full captures, sanitizer
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${tag}_pytest_gpu.txt
for w in config4 config2 config3 readme1024 readme2048; do
  timeout 600 python bench.py --steps 20 --warmup 3 --workload $w > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
for w in config4 config2 config3; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_$w.csv python bench.py --steps 3 --warmup 3 --workload $w --no-cpu-baseline > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'zero_kernel|surface_tri_kernel' --launch-skip 8 -c 2 -f -o gpurun_out/${tag}_config4_full python bench.py --steps 3 --warmup 3 --workload config4 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'solid_tri_kernel|solid_fill_kernel' --launch-skip 8 -c 2 -f -o gpurun_out/${tag}_config3_full python bench.py --steps 3 --warmup 3 --workload config3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:surface_coop_kernel --launch-skip 4 -c 1 -f -o gpurun_out/${tag}_config2_coop python bench.py --steps 3 --warmup 3 --workload config2 --no-cpu-baseline > /dev/null 2>&1
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize.py > gpurun_out/${tag}_sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize.py > gpurun_out/${tag}_sanitizer_racecheck.log 2>&1
cat gpurun_out/${tag}_pytest_gpu.txt; tail -3 gpurun_out/${tag}_sanitizer_memcheck.log; tail -3 gpurun_out/${tag}_sanitizer_racecheck.log
for w in config4 config2 config3 readme1024 readme2048; do python -c "
import json;d=json.load(open('gpurun_out/${tag}_bench_$w.json'));print('$w',d['value'],d['ms_per_step'],d['roofline']['phases_ms'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline'] and d['cpu_baseline']['value'])"; done
cat gpurun_out/${tag}_bench_reference.json
