mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu > gpurun_out/r02e_tests.log 2>&1; tail -5 gpurun_out/r02e_tests.log
B="python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-e2e --others 0"
for v in default v1 persistent h1; do
  if [ $v = default ]; then timeout 300 $B > gpurun_out/r02e_$v.json 2> gpurun_out/r02e_$v.err
  elif [ $v = v1 ]; then SBX_RESIDENT_V1=1 timeout 300 $B > gpurun_out/r02e_$v.json 2> gpurun_out/r02e_$v.err
  else SBX_LIB=$PWD/sbsim_b200/lib/variants/libsbx_$v.so timeout 300 $B > gpurun_out/r02e_$v.json 2> gpurun_out/r02e_$v.err; fi
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
  l=json.load(open(f'gpurun_out/r02e_{v}.json')); r=l['roofline']
  print(v, 'value %.2fM ms/step %.3f solve_ms %.3f frac %.3f sweeps %.2f' % (l['value']/1e6, l['ms_per_step'], r['launch_ms'], r['frac'], r['mean_sweeps_per_step']))
except Exception as e:
  print(v, 'FAILED', e); print(open(f'gpurun_out/r02e_{v}.err').read()[-800:])
PY
done
