# bench.py for the three workloads at N GPUs (default 8): bash tools/bench_all_workloads_n8.sh [N]
N=${1:-8}
mkdir -p gpurun_out
python bench.py --gpus $N --steps 4 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/bench_n${N}_uniform16.json
python bench.py --gpus $N --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --workload mostly_sorted16 2>/dev/null | grep '^{' > gpurun_out/bench_n${N}_mostly_sorted16.json
python bench.py --gpus $N --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --workload particles48 2>/dev/null | grep '^{' > gpurun_out/bench_n${N}_particles48.json
for f in gpurun_out/bench_n${N}_*.json; do python -c "
import sys,json
d=json.load(open('$f')); print(d['config']['workload'][:40], 'ms/step %.2f  %.2f Grec/s  %.1f GB/s'%(d['ms_per_step'], d['value']/1e9, d['gb_per_s'])); print('  phases', [(k,round(v,2)) for k,v in d['phases_ms'] if v>0.2]); print('  exch', d['exchange']); print('  kern', {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}); print('  ', d['local_sort'])
"; done
