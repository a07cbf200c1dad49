mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "legacy" > gpurun_out/r02ae_tests.log 2>&1; tail -5 gpurun_out/r02ae_tests.log
