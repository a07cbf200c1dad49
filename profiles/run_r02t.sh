mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "resident_kernels or fd_step or many_rooms" > gpurun_out/r02t_tests.log 2>&1; tail -3 gpurun_out/r02t_tests.log
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --others 0"
timeout 300 $B > gpurun_out/r02t_v3.json 2> gpurun_out/r02t_v3.err
python - <<'PY'
import json
try:
  l=json.load(open('gpurun_out/r02t_v3.json')); r=l['roofline']
  print('value %.2fM ms/step %.3f solve_ms %.3f frac %.3f sweeps %.2f %s' % (l['value']/1e6, l['ms_per_step'], r['launch_ms'], r['frac'], r['mean_sweeps_per_step'], r['kernel'][:18]))
except Exception as e:
  print('FAILED', e); print(open('gpurun_out/r02t_v3.err').read()[-1500:])
PY
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --others 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_resident_step3" -s 6 -c 1 -o gpurun_out/r02t_v3 -f $B > gpurun_out/r02t_ncu.log 2>&1
ls -la gpurun_out/r02t_v3.ncu-rep
