mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --others 0"
for v in default v1 default v1; do
  if [ $v = default ]; then timeout 300 $B > gpurun_out/r02e_$v.json 2> gpurun_out/r02e_$v.err
  else SBX_RESIDENT_V1=1 timeout 300 $B > gpurun_out/r02e_$v.json 2> gpurun_out/r02e_$v.err; fi
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
l=json.load(open(f'gpurun_out/r02e_{v}.json')); r=l['roofline']
print(v, 'value %.2fM ms/step %.3f solve_ms %.3f frac %.3f sweeps %.2f e2e %.2fM' % (l['value']/1e6, l['ms_per_step'], r['launch_ms'], r['frac'], r['mean_sweeps_per_step'], l['e2e']['value']/1e6))
PY
done
