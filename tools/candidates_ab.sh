#!/bin/bash
# One gpurun call that settles the candidates prepared at the end of round 1 (written without a GPU).
#   bash tools/candidates_ab.sh            (1 GPU: parity of every candidate, merge timing on rank threads)
#   bash tools/candidates_ab.sh 8          (8 GPUs: bench.py A/B of the merge candidate at the headline config)
# Output goes to gpurun_out/candidates_ab.log; copy it to profiles/ once read.
cd "$(dirname "$0")/.."
N=${1:-1}
LOG=gpurun_out/candidates_ab.log
mkdir -p gpurun_out
{
echo "# candidates A/B, $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1), N=$N"
echo "== parity of the candidates (tests/test_zz_candidates.py)"
MPSORT_TEST_CANDIDATES=1 timeout 1800 python -m pytest tests/test_zz_candidates.py -q -m gpu 2>&1 | tail -8
echo "== parity of the callback entry points (tests/test_zz_callback_api.py)"
timeout 900 python -m pytest tests/test_zz_callback_api.py -q -m gpu 2>&1 | tail -3
for P in 2 4 8; do
  echo "== merge kernels, $P rank threads on one GPU, 2^26 16-byte records per rank: rounds | buckets"
  timeout 600 python tools/group_probe.py $P 26 16 0 2>&1 | tail -3
  MPSORT_MERGE_BUCKET=1 timeout 600 python tools/group_probe.py $P 26 16 0 2>&1 | tail -3
done
echo "== mostly sorted keys and 48-byte duplicates-heavy records (fallback share matters here), 8 rank threads"
for K in "16 1" "48 2"; do
  timeout 600 python tools/group_probe.py 8 25 $K 2>&1 | tail -3
  MPSORT_MERGE_BUCKET=1 timeout 600 python tools/group_probe.py 8 25 $K 2>&1 | tail -3
done
if [ "$N" -gt 1 ]; then
  echo "== bench.py --gpus $N: default | bucket merge | peer splitter kernel | both"
  bash tools/exchange_modes.sh $N "MPSORT_MERGE_BUCKET=0" "MPSORT_MERGE_BUCKET=1" "MPSORT_PEER_SPLITTER=1" "MPSORT_PEER_SPLITTER=1 MPSORT_MERGE_BUCKET=1"
  echo "== bench.py --gpus $N --workload mostly_sorted16: default | MPSORT_HYBRID_DEPTH5=1"
  for cfg in "MPSORT_X=0" "MPSORT_HYBRID_DEPTH5=1"; do
    echo "-- $cfg"
    env $cfg python bench.py --gpus $N --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --workload mostly_sorted16 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.2f  value %.2f Grec/s'%(d['ms_per_step'], d['value']/1e9)); print('  phases', [(k,round(v,2)) for k,v in d['phases_ms'] if v>0.2]); print('  kern', {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}, d['local_sort'])
"
  done
  echo "== bench.py --gpus $N --workload particles48: default | pipelined pack (2, 4 parts, + bucket merge) | fused pack + peer stores"
  for cfg in "MPSORT_X=0" "MPSORT_PACK_PIPELINE=1" "MPSORT_PACK_PIPELINE=1 MPSORT_EXCHANGE_PHASES=4" "MPSORT_PACK_PIPELINE=1 MPSORT_MERGE_BUCKET=1" \
             "MPSORT_FUSED_PACK=1 MPSORT_P2P_CE=0 MPSORT_EXCHANGE_PHASES=1" "MPSORT_FUSED_PACK=1 MPSORT_P2P_CE=0 MPSORT_EXCHANGE_PHASES=1 MPSORT_P2P_CTAS_PER_SM=1"; do
    echo "-- $cfg"
    env $cfg python bench.py --gpus $N --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --workload particles48 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.2f  value %.2f Grec/s'%(d['ms_per_step'], d['value']/1e9)); print('  phases', [(k,round(v,2)) for k,v in d['phases_ms'] if v>0.2]); print('  kern', {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
"
  done
fi
if [ -x mp-sort_b200/variants/libmpsort-b200.persist.so ] || bash tools/build_variant.sh persist "-DMPSK_REC_PERSIST=1" >/dev/null 2>&1; then
  echo "== record pass: default | persistent v2 (MPSK_REC_PERSIST=1)"
  python tools/sweep.py 28 16 0
  MPSORT_LIB=$PWD/mp-sort_b200/variants/libmpsort-b200.persist.so python tools/sweep.py 28 16 0
fi
} 2>&1 | tee $LOG
