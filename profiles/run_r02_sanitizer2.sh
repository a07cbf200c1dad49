mkdir -p gpurun_out
export PYTHONPATH=$PWD
# the kernels added after the first sanitizer pass: k_pw_leaves / k_pw_combine, k_resident_step<4, true>,
# the first / later k_sweep instantiations inside the graph loop, large-grid Gauss-Seidel, k_convect_reduce
SEL="means_are_np_mean or shared_plans or fd_step_bit_exact or device_rng_convection_is_the_same or larger_than_shared"
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_numpy_means.py tests/test_gpu_parity.py tests/test_convection.py -x -q -m gpu -k "$SEL" > gpurun_out/r02_sanitizer2_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r02_sanitizer2_$tool.log | tail -4
done
