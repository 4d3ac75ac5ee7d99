#!/bin/bash
# usage: scripts/gpurun_bench.sh <tag> [extra bench args]  — short bench of the headline config + parity check
tag=$1; shift
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$tag.json"))
    print("value",d["value"],"ms",d["ms_per_step"],d["roofline"]["phases_ms"],"parity",d["parity"],"e2e",d["e2e"]["value"], d["counters"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_$tag.err").read()[-2000:])
PY
