mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
  l=json.load(open(f'gpurun_out/r02_bench_n{n}.json')); r=l['roofline']
  print('N=%s value %.2fM ms %.3f e2e %.2fM frac %.3f' % (n, l['value']/1e6, l['ms_per_step'], l['e2e']['value']/1e6, r['frac']))
  print('with_allgather', {k:(round(v,3) if isinstance(v,float) else v) for k,v in l.get('with_allgather',{}).items() if k!='collective'})
  for o in l.get('other_configs', []):
    print(o.get('error') or (o['config']['workload'][:30], 'value %.2fM e2e %.2fM' % (o['value']/1e6, o['e2e']['value']/1e6), o.get('with_allgather',{}).get('ms_per_step')))
except Exception as e:
  print('FAILED', e); print(open(f'gpurun_out/r02_bench_n{n}.err').read()[-1500:])
PY
