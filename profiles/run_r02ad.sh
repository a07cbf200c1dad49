mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_convect_reduce|k_zone_reduce" -c 6 --csv --log-file gpurun_out/r02ad.csv python bench.py --workload office --envs-per-gpu 512 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02ad.log 2>&1
grep -E "k_convect|k_zone" gpurun_out/r02ad.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | head -40
