#!/bin/bash
# after k_sweep_tma became the default sweep: the office captures and the bench lines again
mkdir -p gpurun_out
bash profiles/capture_r02_office.sh
timeout 1500 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_steps20.json 2> gpurun_out/r02_bench_steps20.err
cut -c1-200 gpurun_out/r02_bench.json
