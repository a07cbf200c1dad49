#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bounds or device_pointer or auto_reset or episode" > gpurun_out/r02e2e_tests.log 2>&1; tail -3 gpurun_out/r02e2e_tests.log
timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --others 0 > gpurun_out/r02e2e.json 2> gpurun_out/r02e2e.err
python - <<PY
import json
l=json.load(open("gpurun_out/r02e2e.json"))
print("value", round(l["value"]/1e6,2), "e2e", round(l["e2e"]["value"]/1e6,2), "ratio", round(l["e2e"]["value"]/l["value"],3))
PY
