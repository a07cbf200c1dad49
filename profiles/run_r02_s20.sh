mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_steps20.json 2> gpurun_out/r02_bench_steps20.err
python - <<'PY'
import json
l=json.load(open('gpurun_out/r02_bench_steps20.json')); r=l['roofline']
print('value %.2fM ms %.3f e2e %.2fM frac %.3f whole %.3f sweeps %.2f launch_ms %.4f clocks %s' % (l['value']/1e6, l['ms_per_step'], l['e2e']['value']/1e6, r['frac'], r['frac_whole_step'], r['mean_sweeps_per_step'], r['launch_ms'], l['clocks']))
for o in l.get('other_configs', []):
  if 'error' in o: print('ERR', o); continue
  print(o['config']['workload'][:40], o['config'].get('stochastic_convection','')[:12], 'value %.1fk ms %.3f e2e %.1fk frac %.3f whole %.3f' % (o['value']/1e3, o['ms_per_step'], o['e2e']['value']/1e3, o['roofline']['frac'], o['roofline']['frac_whole_step']))
PY
