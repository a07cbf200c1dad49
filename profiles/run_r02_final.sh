#!/bin/bash
# whole GPU suite + smoke
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_tests.log 2>&1; tail -4 gpurun_out/r02_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
