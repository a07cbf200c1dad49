mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02k_n2.json 2> gpurun_out/r02k_n2.err
python - <<'PY'
import json
try:
  l=json.load(open('gpurun_out/r02k_n2.json')); r=l['roofline']
  print('N=2 value %.2fM e2e %.2fM frac %.3f' % (l['value']/1e6, l['e2e']['value']/1e6, r['frac']))
  print('with_allgather', {k:(round(v,3) if isinstance(v,float) else v) for k,v in l['with_allgather'].items() if k!='collective'})
  for o in l.get('other_configs', []):
    print(o.get('error') or (o['config']['workload'][:30], 'value %.2fM e2e %.2fM' % (o['value']/1e6, o['e2e']['value']/1e6), o.get('with_allgather',{}).get('ms_per_step')))
except Exception as e:
  print('FAILED', e); print(open('gpurun_out/r02k_n2.err').read()[-1500:])
PY
