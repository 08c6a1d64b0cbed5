#!/bin/bash
# round 2, GPU call 1 (1 GPU): state of the tree on hardware + the candidates that were never measured
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -1); $(nproc) host threads"
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== merge alone, one GPU, full size: rounds (default) | buckets"
for P in 8 4 2; do
  timeout 300 python tools/merge_probe.py $P 28 16 0 5 2>&1 | tail -1
  MPSORT_MERGE_BUCKET=1 timeout 300 python tools/merge_probe.py $P 28 16 0 5 2>&1 | tail -1
done
timeout 300 python tools/merge_probe.py 8 28 16 1 3 2>&1 | tail -1
MPSORT_MERGE_BUCKET=1 timeout 300 python tools/merge_probe.py 8 28 16 1 3 2>&1 | tail -1
timeout 300 python tools/merge_probe.py 8 27 48 2 3 2>&1 | tail -1
MPSORT_MERGE_BUCKET=1 timeout 300 python tools/merge_probe.py 8 27 48 2 3 2>&1 | tail -1
echo "== record pass variants (tools/sweep.py 28 16 0): default | nobulk+static | persist | static"
timeout 600 bash tools/run_sweep.sh 28 16 0 2>&1 | tail -5
echo "== mostly sorted: default | MPSORT_HYBRID_DEPTH5=1"
timeout 200 python tools/sweep.py 28 16 1 2>&1 | tail -1
MPSORT_HYBRID_DEPTH5=1 timeout 200 python tools/sweep.py 28 16 1 2>&1 | tail -1
echo "== bench.py N=1"
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -c 600 gpurun_out/bench_n1.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_n1.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value %.2f Grec/s  ms %.3f  e2e %.1f ms  roofline %.3f  preflight %s (%.1f s)' % (d['value']/1e9, d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['parity_preflight']['ok'], d['parity_preflight']['seconds']))
        print(' cpu', d['cpu_baseline'])
        print(' kernels', {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
        for w,v in d['workloads'].items():
            print(' ', w, '%.2f Grec/s %.2f ms' % (v['value']/1e9, v['ms_per_step']), {k: round(x['ms_per_step'],2) for k,x in v['kernels'].items()}, v['local_sort'])
PY
echo "== reference arm, full size"
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 --ref-budget-s 45 2>&1 | tail -1 | cut -c1-900
echo "== compute-sanitizer (memcheck, racecheck; small inputs)"
TOOLS="memcheck racecheck" SANITIZE_TIMEOUT=200 SEL='test_radix_sort_desc_matches_oracle or test_golden_vectors or test_second_sort_merge_path or test_record_mode_bare_8_byte_keys' bash tools/sanitize.sh 2>&1 | tail -6
} 2>&1 | tee gpurun_out/call1.log
