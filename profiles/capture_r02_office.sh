#!/bin/bash
# k_sweep runs inside a CUDA-graph WHILE node (device-driven sweep loop), which ncu does not
# open; for the profile the same kernels are launched by the host-polled loop
# (SBX_HOST_SWEEP_LOOP=1: identical kernels and arguments, one launch per sweep).
mkdir -p gpurun_out
SBX_HOST_SWEEP_LOOP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_office.csv \
  python bench.py --workload office --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_l2.log 2>&1
SBX_HOST_SWEEP_LOOP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_zone_reduce|k_convect_reduce" -s 8 -c 4 -o gpurun_out/r02_office -f \
  python bench.py --workload office --envs-per-gpu 512 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_n2.log 2>&1
ls -la gpurun_out/r02_office.ncu-rep; grep -c k_sweep gpurun_out/r02_launches_office.csv
