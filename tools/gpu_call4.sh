#!/bin/bash
# round 2, GPU call 4 (1 GPU): full suite after the host-buffer pipeline, merge tile shapes, bench N=1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -1)"
echo "== pytest -m gpu (whole suite)"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== merge alone: default (prologue without the first barrier) and tile shapes"
for P in 8 4; do
  timeout 300 python tools/merge_probe.py $P 28 16 0 5 2>&1 | tail -1
  for V in m2k256 m2k512 m4k256; do MPSORT_LIB=$PWD/mp-sort_b200/variants/libmpsort-b200.$V.so timeout 300 python tools/merge_probe.py $P 28 16 0 5 2>&1 | tail -1; done
done
timeout 300 python tools/merge_probe.py 8 27 48 2 3 2>&1 | tail -1
for V in m2k256 m2k512; do MPSORT_LIB=$PWD/mp-sort_b200/variants/libmpsort-b200.$V.so timeout 300 python tools/merge_probe.py 8 27 48 2 3 2>&1 | tail -1; done
timeout 300 python tools/merge_probe.py 8 25 16 0 5 2>&1 | tail -1
echo "== sorts: uniform16, bare keys, particles48, mostly sorted (own keys, and as rank 7 of 8 holds them: 5 passes expected)"
timeout 150 python tools/sweep.py 28 16 0 | tail -1
timeout 150 python tools/sweep.py 28 8 0 | tail -1
timeout 150 python tools/sweep.py 28 48 2 | tail -1
timeout 150 python tools/sweep.py 28 16 1 | tail -1
SWEEP_AS=7,8 timeout 150 python tools/sweep.py 28 16 1 | tail -1
SWEEP_AS=7,8 MPSORT_NO_HYBRID5=1 timeout 150 python tools/sweep.py 28 16 1 | tail -1
echo "== bench.py N=1 (e2e with chunked host buffers | without)"
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_call4.json 2> gpurun_out/bench_n1_call4.err; echo "rc=$?"; tail -c 400 gpurun_out/bench_n1_call4.err
MPSORT_NO_HOST_CHUNKS=1 timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-preflight --no-extra-workloads > gpurun_out/bench_n1_call4_nochunks.json 2>/dev/null
python - <<'PY'
import json
for f in ('gpurun_out/bench_n1_call4.json', 'gpurun_out/bench_n1_call4_nochunks.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            print(f, 'value %.2f Grec/s  ms %.3f  e2e %.2f ms  roofline %.3f' % (d['value']/1e9, d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac']))
            print(' e2e phases', [(k, round(v,2)) for k,v in d['e2e']['phases_ms']])
            print(' kernels', {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
            for w,v in d.get('workloads', {}).items():
                print(' ', w, '%.2f Grec/s %.2f ms' % (v['value']/1e9, v['ms_per_step']), {k: round(x['ms_per_step'],2) for k,x in v['kernels'].items()})
PY
} 2>&1 | tee gpurun_out/call4.log
