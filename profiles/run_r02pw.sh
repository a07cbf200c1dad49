#!/bin/bash
# SBX_OPT_NUMPY_MEANS: parity and its cost on the headline workload
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_numpy_means.py -m gpu -q > gpurun_out/r02pw_tests.log 2>&1; tail -3 gpurun_out/r02pw_tests.log
for zm in exact-integer numpy; do
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --others 0 --zone-means $zm > gpurun_out/r02pw_bench_$zm.json 2> gpurun_out/r02pw_bench_$zm.err
  python - <<PY
import json
l=json.load(open("gpurun_out/r02pw_bench_$zm.json"))
print("$zm", "value", round(l["value"]/1e6,2), "ms", round(l["ms_per_step"],4), "solve", round(l["roofline"]["launch_ms"],4), "e2e", round(l["e2e"]["value"]/1e6,2), "launches", l["gpu_launches"])
PY
done
