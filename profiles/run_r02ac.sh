mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_convection.py tests/test_gpu_golden.py -x -q -m gpu > gpurun_out/r02ac_tests.log 2>&1; tail -5 gpurun_out/r02ac_tests.log
timeout 600 python bench.py --workload office --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02ac_office.json 2> gpurun_out/r02ac_office.err
python - <<'PY'
import json
try:
  l=json.load(open('gpurun_out/r02ac_office.json')); r=l['roofline']
  print('office value %.1fk ms %.3f e2e %.1fk frac %.3f whole %.3f sweeps %.2f' % (l['value']/1e3, l['ms_per_step'], l['e2e']['value']/1e3, r['frac'], r['frac_whole_step'], r['mean_sweeps_per_step']))
except Exception as e:
  print('FAILED', e); print(open('gpurun_out/r02ac_office.err').read()[-1500:])
PY
