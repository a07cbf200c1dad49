#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for pw in 0 1; do
  PW=$pw SBX_LIB=$PWD/sbsim_b200/lib/libsbx_prof.so timeout 300 python profiles/phase_profile.py > gpurun_out/r02pw_phase_$pw.txt 2>&1
  head -12 gpurun_out/r02pw_phase_$pw.txt
done
