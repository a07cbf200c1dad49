mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --others 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_resident_step3" -s 6 -c 1 -o gpurun_out/r02r_v3 -f $B > gpurun_out/r02r_ncu.log 2>&1
ls -la gpurun_out/r02r_v3.ncu-rep
