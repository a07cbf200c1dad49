mkdir -p gpurun_out
export PYTHONPATH=$PWD
# k_sweep_tma (the default streaming sweep) on the sweep-kernel test: memcheck as the library runs it (graph
# loop); racecheck with the host-polled loop (racecheck crashes on the host side when the tensor-map kernels sit
# inside the conditional graph node)
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streaming_sweep_kernels" > gpurun_out/r02_sanitizer3_memcheck.log 2>&1
echo "== memcheck"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_sanitizer3_memcheck.log | tail -3
SBX_HOST_SWEEP_LOOP=1 timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streaming_sweep_kernels" > gpurun_out/r02_sanitizer3_racecheck.log 2>&1
echo "== racecheck (host loop)"; grep -E "RACECHECK SUMMARY|passed|failed|Error" gpurun_out/r02_sanitizer3_racecheck.log | tail -4
