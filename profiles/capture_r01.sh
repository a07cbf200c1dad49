#!/bin/bash
# Round-1 evidence run (one B200, via gpurun): bench line, ncu launch lists of the same
# command, one `ncu --set full` capture per hot kernel.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err
python bench.py --impl reference --steps 24 --warmup 3 > gpurun_out/r01_bench_reference.json 2> gpurun_out/r01_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_randomized.csv \
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --others 0 > gpurun_out/r01_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_launches_office.csv \
  python bench.py --workload office --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r01_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_resident_step|k_pre|k_post" -s 9 -c 3 -o gpurun_out/r01_randomized -f \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --others 0 > gpurun_out/r01_n1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_zone_reduce" -s 4 -c 3 -o gpurun_out/r01_office -f \
  python bench.py --workload office --envs-per-gpu 512 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r01_n2.log 2>&1
cut -c1-600 gpurun_out/r01_bench.json
