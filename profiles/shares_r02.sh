mkdir -p gpurun_out
for n in 1 2 4 8; do timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --others 0 --host-shares $n > gpurun_out/r02h_$n.json 2> gpurun_out/r02h_$n.err; python -c "
import json; l=json.load(open('gpurun_out/r02h_$n.json')); print('shares $n: device %.3f ms e2e %.3f ms (%.2fM)' % (l['ms_per_step'], l['e2e']['ms_per_step'], l['e2e']['value']/1e6))"; done
