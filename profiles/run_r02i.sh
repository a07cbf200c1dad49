mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_convection.py -x -q -m gpu > gpurun_out/r02i_tests.log 2>&1; tail -4 gpurun_out/r02i_tests.log
timeout 200 python bench.py --workload office --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_office_conv.json 2> gpurun_out/r02i_office_conv.err
python - <<'PY'
import json
for n in ('office_conv',):
  try:
    l=json.load(open(f'gpurun_out/r02i_{n}.json')); r=l['roofline']
    print(n,'value %.1fk ms/step %.3f e2e %.1fk (%.3f ms) solve_ms %.3f frac %.3f whole %.3f sweeps %.2f launches %d' % (l['value']/1e3, l['ms_per_step'], l['e2e']['value']/1e3, l['e2e']['ms_per_step'], r['launch_ms'], r['frac'], r['frac_whole_step'], r['mean_sweeps_per_step'], l['gpu_launches']))
  except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02i_{n}.err').read()[-600:])
PY
