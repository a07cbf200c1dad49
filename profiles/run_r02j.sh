mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_golden.py -x -q -m gpu -k "one_day" > gpurun_out/r02j_tests.log 2>&1; tail -3 gpurun_out/r02j_tests.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err ) 2>&1 | grep real
python - <<'PY'
import json
l=json.load(open('gpurun_out/r02j_bench.json')); r=l['roofline']
print('main value %.2fM e2e %.2fM frac %.3f sweeps %.2f clocks %s cpu %.0f' % (l['value']/1e6, l['e2e']['value']/1e6, r['frac'], r['mean_sweeps_per_step'], l['clocks'], l['cpu_baseline']['value']))
for o in l.get('other_configs', []):
  if 'error' in o: print('ERR', o); continue
  print(o['config']['workload'][:40], o['config'].get('stochastic_convection','')[:12], 'value %.1fk ms %.3f e2e %.1fk frac %.3f whole %.3f cpu %s' % (o['value']/1e3, o['ms_per_step'], o['e2e']['value']/1e3, o['roofline']['frac'], o['roofline']['frac_whole_step'], o.get('cpu_baseline',{}).get('value')))
PY
