#!/bin/bash
# round 2, 8-GPU call: bench.py --gpus 8 (pre-flight through the production transport, headline, e2e, configs 4/5),
# then every candidate switch inside ONE launch (tools/ab_multi.py), NCCL parity worker, topology
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c | tr '\n' ' '); $(nproc) host threads; $(lscpu | grep -i 'numa node(s)')"
nvidia-smi topo -m 2>/dev/null | head -12
echo "== bench.py --gpus 8"
timeout 1200 python bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_n8.err | tail -c 1500
python - <<'PY'
import json
for l in open('gpurun_out/bench_n8.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value %.2f Grec/s  ms %.3f  e2e %.1f ms = %.2f Grec/s (%s)  roofline %.3f' % (d['value']/1e9, d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value']/1e9, d['e2e'].get('numa'), d['roofline']['frac']))
        pf=d['parity_preflight']; print(' preflight', pf['ok'], pf['seconds'], 's', pf['transport'], [(c['workload'], c['exchange_parts'], c['all_ranks_ok']) for c in pf['cases']])
        print(' phases', [(k,round(v,2)) for k,v in d['phases_ms'] if v>0.1]); print(' exchange', d['exchange'])
        print(' kernels', {k: (round(v['ms_per_step'],3), round(v.get('frac_of_peak',0),3)) for k,v in d['kernels'].items()})
        print(' e2e phases', [(k, round(v,2)) for k,v in d['e2e']['phases_ms']])
        for w,v in d['workloads'].items():
            print(' ', w, '%.2f Grec/s %.2f ms' % (v['value']/1e9, v['ms_per_step']), [(k,round(x,2)) for k,x in v['phases_ms'] if x>0.1], {k: round(x['ms_per_step'],2) for k,x in v['kernels'].items()}, v['exchange'])
PY
echo "== switches, one launch"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/ab_multi.py --gpus 8 --steps 5 \
  "uniform16:-" "uniform16:MPSORT_EXCHANGE_PHASES=3" "uniform16:MPSORT_EXCHANGE_PHASES=4" "uniform16:MPSORT_EXCHANGE_PHASES=8" \
  "uniform16:MPSORT_PEER_SPLITTER=1" "uniform16:MPSORT_PEER_SPLITTER=1,MPSORT_EXCHANGE_PHASES=4" "uniform16:MPSORT_EXCHANGE_PHASES=1" \
  "particles48:-" "particles48:MPSORT_EXCHANGE_PHASES=4" "particles48:MPSORT_PACK_PIPELINE=1" "particles48:MPSORT_PACK_PIPELINE=1,MPSORT_EXCHANGE_PHASES=4" \
  "particles48:MPSORT_FUSED_PACK=1" "particles48:MPSORT_FUSED_PACK=1,MPSORT_EXCHANGE_PHASES=1" "particles48:MPSORT_FUSED_PACK=1,MPSORT_EXCHANGE_PHASES=4" \
  "mostly_sorted16:-" "mostly_sorted16:MPSORT_EXCHANGE_PHASES=1" "mostly_sorted16:MPSORT_NO_HYBRID5=1" 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$"
echo "== NCCL process-per-GPU parity at 8 ranks (tests/nccl_worker.py)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "nccl_process" 2>&1 | tail -3
} 2>&1 | tee gpurun_out/call_n8.log
