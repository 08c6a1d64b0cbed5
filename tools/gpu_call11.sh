#!/bin/bash
# round 2, last 1-GPU call: what the driver runs at round end, on the final code -- smoke(), bench.py with its defaults
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | tr '\n' ' ')"
echo "== smoke()"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== python bench.py"
SECONDS=0; timeout 1200 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "rc=$?"
echo "wall ${SECONDS} s"
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_n1_final.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value %.2f Grec/s  ms %.3f  steps %d warmup %d  roofline %s' % (d['value']/1e9, d['ms_per_step'], d['steps'], d['warmup'], {k: d['roofline'][k] for k in ('achieved','peak','frac','traffic')}))
        e=d['e2e']; print(' e2e %.1f ms = %.2f Grec/s' % (e['ms_per_step'], e['value']/1e9)); print(' cpu_baseline', d['cpu_baseline']); print(' clocks', d['clocks']); print(' launches', d['gpu_launches'])
        pf=d['parity_preflight']; print(' preflight', pf['ok'], pf['seconds'])
        for w,v in d['workloads'].items(): print(' ', w, '%.2f Grec/s %.2f ms' % (v['value']/1e9, v['ms_per_step']), {k: round(x['ms_per_step'],2) for k,x in v['kernels'].items()})
PY
} 2>&1 | tee gpurun_out/call11.log
