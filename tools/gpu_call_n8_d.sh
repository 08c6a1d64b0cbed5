#!/bin/bash
# round 2, last 8-GPU call: own slice merged in place, gate modes of the exchange parts (A/B inside one launch), then bench.py at 8, 4 and 2 GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l)
        e=d.get('e2e') or {}
        print('value %.2f Grec/s  ms %.3f  roofline %.3f  %s' % (d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['transport']))
        if e: print(' e2e %.1f ms = %.2f Grec/s' % (e['ms_per_step'], e['value']/1e9), [(k, round(v,1)) for k,v in e['phases_ms']])
        pf=d['parity_preflight']; print(' preflight', pf['ok'], pf['seconds'], 's', [(c['workload'], c['exchange_parts'], c['all_ranks_ok']) for c in pf['cases']])
        print(' phases', [(k,round(v,2)) for k,v in d['phases_ms'] if v>0.1]); print(' exchange', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['exchange'].items()})
        print(' kernels', {k: (round(v['ms_per_step'],3), round(v.get('frac_of_peak',0),3)) for k,v in d['kernels'].items()})
        for w,v in d['workloads'].items():
            print(' ', w, '%.2f Grec/s %.2f ms' % (v['value']/1e9, v['ms_per_step']), [(k,round(x,2)) for k,x in v['phases_ms'] if x>0.1], {k: round(x['ms_per_step'],2) for k,x in v['kernels'].items()}, {k:(round(x,3) if isinstance(x,float) else x) for k,x in v['exchange'].items()})
PY
}
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c | tr '\n' ' '); $(nproc) host threads"
echo "== one launch: defaults / old part order / more parts / the other two workloads"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 tools/ab_multi.py --gpus 8 --steps 8 \
  "uniform16:-" "uniform16:MPSORT_CHAINED_PARTS=2" "uniform16:MPSORT_NO_SELF_IN_PLACE=1" "uniform16:-" \
  "mostly_sorted16:-" "mostly_sorted16:MPSORT_NO_SELF_IN_PLACE=1" "mostly_sorted16:MPSORT_EXCHANGE_PHASES=1" 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$"
echo "== bench.py --gpus 8"
timeout 1200 python bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_n8_d.json 2> gpurun_out/r02_bench_n8_d.err; echo "rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/r02_bench_n8_d.err | tail -c 1000
show gpurun_out/r02_bench_n8_d.json
echo "== bench.py --gpus 4"
timeout 900 python bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/r02_bench_n4_d.json 2> gpurun_out/r02_bench_n4_d.err; echo "rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/r02_bench_n4_d.err | tail -c 1000
show gpurun_out/r02_bench_n4_d.json
echo "== bench.py --gpus 2"
timeout 900 python bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/r02_bench_n2_d.json 2> gpurun_out/r02_bench_n2_d.err; echo "rc=$?"
show gpurun_out/r02_bench_n2_d.json
} 2>&1 | tee gpurun_out/call_n8_d.log
