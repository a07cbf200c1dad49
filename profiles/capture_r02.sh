#!/bin/bash
# Round-2 evidence run (one B200, via gpurun): GPU test suite, smoke, bench lines, ncu launch
# lists of the same commands, one `ncu --set full` capture per hot kernel.  Outputs land in
# gpurun_out/; profiles/summarize_r02.py turns them into the tracked summaries.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gputests.log 2>&1; tail -3 gpurun_out/r02_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 900 python bench.py --impl reference --steps 24 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_randomized.csv \
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --others 0 > gpurun_out/r02_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_office.csv \
  python bench.py --workload office --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_l2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_resident_step|k_pre|k_post" -s 9 -c 3 -o gpurun_out/r02_randomized -f \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --others 0 > gpurun_out/r02_n1.log 2>&1
SBX_RESIDENT_V3=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_resident_step3" -s 3 -c 1 -o gpurun_out/r02_randomized_v3 -f \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --others 0 > gpurun_out/r02_n3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_zone_reduce|k_convect_reduce" -s 4 -c 3 -o gpurun_out/r02_office -f \
  python bench.py --workload office --envs-per-gpu 512 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_n2.log 2>&1
ls -la gpurun_out/r02_*.ncu-rep
cut -c1-400 gpurun_out/r02_bench.json
