mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q -m gpu > gpurun_out/r02ab_tests.log 2>&1; tail -3 gpurun_out/r02ab_tests.log
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --others 0"
timeout 300 $B > gpurun_out/r02ab.json 2> gpurun_out/r02ab.err
python - <<'PY'
import json
try:
  l=json.load(open('gpurun_out/r02ab.json')); r=l['roofline']
  print('value %.2fM ms/step %.3f solve_ms %.3f frac %.3f sweeps %.2f %s' % (l['value']/1e6, l['ms_per_step'], r['launch_ms'], r['frac'], r['mean_sweeps_per_step'], r['kernel'][:18]))
except Exception as e:
  print('FAILED', e); print(open('gpurun_out/r02ab.err').read()[-1500:])
PY
