mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02g_tests.log 2>&1; tail -5 gpurun_out/r02g_tests.log
timeout 600 python bench.py --workload office --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_office.json 2> gpurun_out/r02g_office.err
SBX_HOST_SWEEP_LOOP=1 timeout 600 python bench.py --workload office --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_office_host.json 2> gpurun_out/r02g_office_host.err
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --others 0 > gpurun_out/r02g_rand.json 2> gpurun_out/r02g_rand.err
python - <<'PY'
import json
for n in ('office','office_host','rand'):
  try:
    l=json.load(open(f'gpurun_out/r02g_{n}.json')); r=l['roofline']
    print(n,'value %.1fk ms/step %.3f e2e %.1fk (%.3f ms) solve_ms %.3f frac %.3f whole %.3f sweeps %.2f launches %d' % (l['value']/1e3, l['ms_per_step'], l['e2e']['value']/1e3, l['e2e']['ms_per_step'], r['launch_ms'], r['frac'], r['frac_whole_step'], r['mean_sweeps_per_step'], l['gpu_launches']))
  except Exception as e:
    print(n,'FAILED',e); print(open(f'gpurun_out/r02g_{n}.err').read()[-600:])
PY
