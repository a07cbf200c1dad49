mkdir -p gpurun_out
export PYTHONPATH=$PWD
# k_sweep_tma (the default streaming sweep): both tools on the sweep-kernel test
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streaming_sweep_kernels" > gpurun_out/r02_sanitizer3_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r02_sanitizer3_$tool.log | tail -3
done
