mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "resident_kernels or fd_step or many_rooms" > gpurun_out/r02w_tests.log 2>&1; tail -5 gpurun_out/r02w_tests.log
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --others 0"
for v in 3; do
  SBX_RESIDENT_V3=$v timeout 300 $B > gpurun_out/r02w_v$v.json 2> gpurun_out/r02w_v$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
  l=json.load(open(f'gpurun_out/r02w_v{v}.json')); r=l['roofline']
  print('V3=%s value %.2fM ms/step %.3f solve_ms %.3f frac %.3f sweeps %.2f %s' % (v, l['value']/1e6, l['ms_per_step'], r['launch_ms'], r['frac'], r['mean_sweeps_per_step'], r['kernel'][:18]))
except Exception as e:
  print('FAILED', e); print(open(f'gpurun_out/r02w_v{v}.err').read()[-1500:])
PY
done
SBX_LIB=$PWD/sbsim_b200/lib/variants/libsbx_phases3.so timeout 300 python profiles/phase_profile3.py 2>&1 | tail -12
