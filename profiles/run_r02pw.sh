#!/bin/bash
# SBX_OPT_NUMPY_MEANS: parity and its cost on the headline and office workloads
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_numpy_means.py -m gpu -x -q > gpurun_out/r02pw_tests.log 2>&1; tail -3 gpurun_out/r02pw_tests.log
for zm in exact-integer numpy; do
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --others 0 --zone-means $zm > gpurun_out/r02pw_bench_$zm.json 2> gpurun_out/r02pw_bench_$zm.err
  python - <<PY
import json
l=json.load(open("gpurun_out/r02pw_bench_$zm.json"))
print("$zm", "value", round(l["value"]/1e6,2), "ms", round(l["ms_per_step"],4), "solve", round(l["roofline"]["launch_ms"],4), "e2e", round(l["e2e"]["value"]/1e6,2), "launches", l["gpu_launches"])
PY
  for conv in off default; do
  timeout 300 python bench.py --workload office --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --others 0 --convection $conv --zone-means $zm > gpurun_out/r02pw_office_${conv}_$zm.json 2> gpurun_out/r02pw_office_${conv}_$zm.err
  python - <<PY
import json
l=json.load(open("gpurun_out/r02pw_office_${conv}_$zm.json"))
print("office conv=$conv $zm", "value", round(l["value"]/1e3,1), "k ms", round(l["ms_per_step"],3), "launches", l["gpu_launches"])
PY
  done
done
