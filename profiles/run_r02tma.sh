#!/bin/bash
# TMA-staged streaming sweep (SBX_SWEEP_TMA=1): parity, then A/B on the office workloads
mkdir -p gpurun_out
export PYTHONPATH=$PWD
SBX_SWEEP_TMA=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q -k "streaming or calibrated or fd_step" > gpurun_out/r02tma_tests.log 2>&1; tail -4 gpurun_out/r02tma_tests.log
for tma in 0 1; do
  if [ $tma = 1 ]; then export SBX_SWEEP_TMA=1; else unset SBX_SWEEP_TMA; fi
  for conv in default off; do
  timeout 300 python bench.py --workload office --steps 12 --warmup 3 --no-cpu-baseline --no-e2e --others 0 --convection $conv > gpurun_out/r02tma_$tma.json 2> gpurun_out/r02tma_$tma.err
  python - <<PY
import json
try:
  l=json.load(open("gpurun_out/r02tma_$tma.json")); r=l["roofline"]
  print("tma=$tma conv=$conv", "ms", round(l["ms_per_step"],3), "sweeps ms", round(r["launch_ms"],3), "frac", round(r["frac"],3))
except Exception as e:
  print("FAILED", e); print(open("gpurun_out/r02tma_$tma.err").read()[-800:])
PY
  done
done
