mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "resident_kernels or fd_step or many_rooms" > gpurun_out/r02n_tests.log 2>&1; tail -15 gpurun_out/r02n_tests.log
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --others 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_resident_step3" -s 6 -c 1 -o gpurun_out/r02n_v3 -f $B > gpurun_out/r02n_ncu.log 2>&1
ls -la gpurun_out/r02n_v3.ncu-rep
