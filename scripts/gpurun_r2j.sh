#!/bin/bash
# host-side bandwidth of the box + current numbers for config 3 + launch list of the prepare path
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" ; free -g | head -2
./scripts/micro/host_bw 1024 1 2>&1 | tee gpurun_out/r2j_host_bw.log
timeout 600 python bench.py --steps 20 --warmup 5 --workload config3 --no-cpu-baseline > gpurun_out/r2j_bench_config3.json 2> gpurun_out/r2j_bench_config3.err
tail -3 gpurun_out/r2j_bench_config3.err; cat gpurun_out/r2j_bench_config3.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2j_launches_prep.csv python scripts/prof_tiles.py > gpurun_out/r2j_prof.log 2>&1
grep -v "^==" gpurun_out/r2j_launches_prep.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['Kernel Name'][:60], r['Metric Value'], r['Metric Unit'])
" | head -40
