#!/bin/bash
# round 2: bench.py at four GPUs with eight exchange parts as the default from four ranks on
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c | tr '\n' ' ')"
timeout 900 python bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/r02_bench_n4_e.json 2> gpurun_out/r02_bench_n4_e.err; echo "rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/r02_bench_n4_e.err | tail -c 600
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_n4_e.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value %.2f Grec/s  ms %.3f  %s' % (d['value']/1e9, d['ms_per_step'], d['transport']))
        pf=d['parity_preflight']; print(' preflight', pf['ok'], pf['seconds'], 's', [(c['workload'], c['exchange_parts'], c['all_ranks_ok'], c.get('own_slices_merged_in_place')) for c in pf['cases']])
        print(' phases', [(k,round(v,2)) for k,v in d['phases_ms'] if v>0.1]); print(' exchange', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['exchange'].items()})
        for w,v in d['workloads'].items():
            print(' ', w, '%.2f Grec/s %.2f ms' % (v['value']/1e9, v['ms_per_step']), [(k,round(x,2)) for k,x in v['phases_ms'] if x>0.1])
PY
} 2>&1 | tee gpurun_out/call_n4b.log
