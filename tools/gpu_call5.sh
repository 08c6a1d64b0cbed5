#!/bin/bash
# round 2, GPU call 5 (1 GPU): whole suite with the new defaults, record-pass variants, bench N=1, ncu evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -1)"
echo "== pytest -m gpu (whole suite)"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== record pass variants: default | plain stores | 384x6 4 CTAs/SM | 256x8 5 CTAs/SM"
timeout 600 bash tools/run_sweep.sh 28 16 0 2>&1 | tail -4
echo "== bench.py N=1"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "rc=$?"; tail -c 300 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_n1.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value %.2f Grec/s  ms %.3f  e2e %.2f ms  roofline %.3f' % (d['value']/1e9, d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac']))
        print(' cpu', d['cpu_baseline']['value'], d['cpu_baseline']['seconds'])
        print(' kernels', {k: (round(v['ms_per_step'],3), round(v.get('frac_of_peak',0),3)) for k,v in d['kernels'].items()})
        for w,v in d.get('workloads', {}).items():
            print(' ', w, '%.2f Grec/s %.2f ms' % (v['value']/1e9, v['ms_per_step']), {k: round(x['ms_per_step'],2) for k,x in v['kernels'].items()})
PY
echo "== ncu launch list of one bench step (shares of a step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_n1.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-preflight --no-extra-workloads > /dev/null 2>&1; echo "rc=$? $(wc -l < gpurun_out/r02_launches_n1.csv) lines"
echo "== ncu --set full: record pass, fix-up, histogram pass (tools/sweep.py), merge tiles at 8 runs, 48-byte gather + index pass"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onesweep_rec_kernel -s 5 -c 1 -o gpurun_out/r02_ncu_rec16 -f python tools/sweep.py 28 16 0 > /dev/null 2>&1; echo "rec rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fixup_rec_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_fixup -f python tools/sweep.py 28 16 0 > /dev/null 2>&1; echo "fixup rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:rec_hist_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_hist -f python tools/sweep.py 28 16 0 > /dev/null 2>&1; echo "hist rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:merge_tile_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_merge_p8 -f python tools/merge_probe.py 8 28 16 0 2 > /dev/null 2>&1; echo "merge rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:gather_records_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_gather48 -f python tools/sweep.py 27 48 2 > /dev/null 2>&1; echo "gather rc=$?"
ls -la gpurun_out/*.ncu-rep
} 2>&1 | tee gpurun_out/call5.log
