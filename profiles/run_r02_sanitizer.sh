mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "resident_kernels or fd_step" > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r02_sanitizer_$tool.log | tail -4
done
