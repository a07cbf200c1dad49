#!/bin/bash
# Round-2 final evidence run (one B200, via gpurun): bench lines (default window, the driver's
# 20-step window, the CPU arm, SBX_OPT_NUMPY_MEANS), ncu launch lists of the same commands, one
# `ncu --set full` capture per hot kernel.  Outputs land in gpurun_out/;
# profiles/summarize_r02.py turns them into the tracked summaries.
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_steps20.json 2> gpurun_out/r02_bench_steps20.err
timeout 600 python bench.py --zone-means numpy --no-cpu-baseline --others 0 > gpurun_out/r02_bench_numpy_means.json 2> gpurun_out/r02_bench_numpy_means.err
timeout 900 python bench.py --impl reference --steps 24 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_randomized.csv \
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --others 0 > gpurun_out/r02_l1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_resident_step|k_pre|k_post" -s 9 -c 3 -o gpurun_out/r02_randomized -f \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --others 0 > gpurun_out/r02_n1.log 2>&1
SBX_RESIDENT_V3=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_resident_step3" -s 3 -c 1 -o gpurun_out/r02_randomized_v3 -f \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --others 0 > gpurun_out/r02_n3.log 2>&1
bash profiles/capture_r02_office.sh
ls -la gpurun_out/r02_*.ncu-rep
cut -c1-300 gpurun_out/r02_bench.json
