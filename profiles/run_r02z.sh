mkdir -p gpurun_out
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --others 0"
for t in 768 640; do
  SBX_RESIDENT_V3=1 SBX_R3_THREADS=$t SBX_LIB=$PWD/sbsim_b200/lib/variants/libsbx_r3_768.so timeout 300 $B > gpurun_out/r02z_$t.json 2> gpurun_out/r02z_$t.err
  python - $t <<'PY'
import json,sys
t=sys.argv[1]
try:
  l=json.load(open(f'gpurun_out/r02z_{t}.json')); r=l['roofline']
  print('threads %s value %.2fM solve_ms %.3f frac %.3f sweeps %.2f %s' % (t, l['value']/1e6, r['launch_ms'], r['frac'], r['mean_sweeps_per_step'], r['kernel'][:18]))
except Exception as e:
  print('FAILED', e); print(open(f'gpurun_out/r02z_{t}.err').read()[-800:])
PY
done
