#!/bin/bash
# round 2, 2-GPU call: the production transport (NCCL ranks, IPC-mapped DMA exchange) bit-exact, candidates over
# mapped peer memory, exchange / parts / splitter A-B, bench.py --gpus 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | tr '\n' ' '); $(nproc) host threads"
echo "== NCCL process-per-GPU parity (tests/nccl_worker.py)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "nccl_process" 2>&1 | tail -3
echo "== candidates over NCCL processes + randomised cases as NCCL ranks (opt-in tests)"
MPSORT_TEST_CANDIDATES=1 timeout 1200 python -m pytest tests/test_zz_candidates.py -m gpu -q -k "nccl" 2>&1 | tail -6
echo "== bench.py --gpus 2 (everything: pre-flight, headline, e2e, other workloads)"
timeout 900 python bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -c 800 gpurun_out/bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value %.2f Grec/s  ms %.3f  e2e %.1f ms (%s)  roofline %.3f' % (d['value']/1e9, d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e'].get('numa'), d['roofline']['frac']))
        pf=d['parity_preflight']; print(' preflight', pf['ok'], pf['seconds'], 's', pf['transport'], [(c['workload'], c['exchange_parts'], c['merge_tiles_all_ranks']) for c in pf['cases']])
        print(' phases', [(k,round(v,2)) for k,v in d['phases_ms'] if v>0.1]); print(' exchange', d['exchange'])
        print(' kernels', {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
        print(' e2e phases', [(k, round(v,2)) for k,v in d['e2e']['phases_ms']])
        for w,v in d['workloads'].items():
            print(' ', w, '%.2f Grec/s %.2f ms' % (v['value']/1e9, v['ms_per_step']), {k: round(x['ms_per_step'],2) for k,x in v['kernels'].items()}, v['exchange'])
PY
echo "== uniform16: parts, split copies, peer splitter"
bash tools/bench_ab.sh 2 uniform16 "-" "MPSORT_EXCHANGE_PHASES=1" "MPSORT_EXCHANGE_PHASES=4" "MPSORT_EXCHANGE_PHASES=8" "MPSORT_P2P_SPLIT=2" "MPSORT_P2P_SPLIT=4" "MPSORT_P2P_SPLIT=2 MPSORT_EXCHANGE_PHASES=4" "MPSORT_PEER_SPLITTER=1" "MPSORT_NO_P2P=1"
echo "== particles48: pack pipelined / fused"
bash tools/bench_ab.sh 2 particles48 "-" "MPSORT_PACK_PIPELINE=1" "MPSORT_PACK_PIPELINE=1 MPSORT_EXCHANGE_PHASES=4" "MPSORT_FUSED_PACK=1 MPSORT_P2P_CE=0 MPSORT_EXCHANGE_PHASES=1" "MPSORT_P2P_SPLIT=2"
echo "== mostly_sorted16"
bash tools/bench_ab.sh 2 mostly_sorted16 "-" "MPSORT_EXCHANGE_PHASES=1"
} 2>&1 | tee gpurun_out/call_n2.log
