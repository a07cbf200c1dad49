mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu > gpurun_out/r02x_tests.log 2>&1; tail -4 gpurun_out/r02x_tests.log
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --others 0"
for v in 0 1; do
  SBX_RESIDENT_V3=$v timeout 300 $B > gpurun_out/r02x_v$v.json 2> gpurun_out/r02x_v$v.err
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
  l=json.load(open(f'gpurun_out/r02x_v{v}.json')); r=l['roofline']
  print('V3=%s value %.2fM ms/step %.3f solve_ms %.3f frac %.3f sweeps %.2f %s' % (v, l['value']/1e6, l['ms_per_step'], r['launch_ms'], r['frac'], r['mean_sweeps_per_step'], r['kernel'][:18]))
except Exception as e:
  print('FAILED', e); print(open(f'gpurun_out/r02x_v{v}.err').read()[-1500:])
PY
done
