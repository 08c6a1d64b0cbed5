python bench.py --gpus 8 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/bench_n8_uniform16.json
python bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --workload mostly_sorted16 2>/dev/null | grep '^{' > gpurun_out/bench_n8_mostly_sorted16.json
python bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --workload particles48 2>/dev/null | grep '^{' > gpurun_out/bench_n8_particles48.json
for f in gpurun_out/bench_n8_*.json; do python -c "
import sys,json
d=json.load(open('$f')); print(d['config']['workload'][:40], 'ms/step %.2f  %.2f Grec/s  %.1f GB/s'%(d['ms_per_step'], d['value']/1e9, d['gb_per_s'])); print('  phases', [(k,round(v,2)) for k,v in d['phases_ms'] if v>0.2]); print('  exch', d['exchange']); print('  kern', {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}); print('  ', d['local_sort'])
"; done
