#!/bin/bash
# SBX_OPT_NUMPY_MEANS: parity and its cost on the headline workload
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_numpy_means.py -m gpu -q > gpurun_out/r02pw_tests.log 2>&1; tail -12 gpurun_out/r02pw_tests.log
