mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_golden.py -x -q -m gpu -k "gauss_seidel" ) > gpurun_out/r02af_tests.log 2>&1; tail -12 gpurun_out/r02af_tests.log
