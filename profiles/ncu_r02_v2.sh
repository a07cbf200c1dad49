mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --others 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_resident_step2|k_pre|k_post" -s 24 -c 3 -o gpurun_out/r02_randomized -f $B > gpurun_out/r02_n1.log 2>&1
ls -la gpurun_out/r02_randomized.ncu-rep
