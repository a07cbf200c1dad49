mkdir -p gpurun_out
echo "== phases default (512,h2)"; SBX_LIB=$PWD/sbsim_b200/lib/variants/libsbx_phases.so timeout 300 python profiles/phase_profile2.py 2>&1 | tail -14
echo "== phases t384"; SBX_LIB=$PWD/sbsim_b200/lib/variants/libsbx_phases_t384.so timeout 300 python profiles/phase_profile2.py 2>&1 | tail -14
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --others 0"
M="--metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_resident_step -c 12 --csv"
timeout 300 ncu $M --log-file gpurun_out/r02d_v2.csv $B > /dev/null 2>&1
SBX_RESIDENT_V1=1 timeout 300 ncu $M --log-file gpurun_out/r02d_v1.csv $B > /dev/null 2>&1
SBX_LIB=$PWD/sbsim_b200/lib/variants/libsbx_t384_h2.so timeout 300 ncu $M --log-file gpurun_out/r02d_t384.csv $B > /dev/null 2>&1
for f in v1 v2 t384; do echo == $f; grep -E "k_resident" gpurun_out/r02d_$f.csv | awk -F'","' '{print $1, $(NF-2), $(NF)}' | tail -9; done
