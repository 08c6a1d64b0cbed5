#!/bin/bash
# A/B of run-time switches through bench.py: bash tools/bench_ab.sh NGPUS WORKLOAD "ENV=.. ENV=.." ["ENV=.." ...]
# (a mode is a quoted list of environment assignments; "-" = the defaults). One line of phases / kernels per mode.
N=$1; W=$2; shift 2
for cfg in "$@"; do
  echo "== $W  $cfg"
  E=$cfg; [ "$cfg" = "-" ] && E="MPSORT_AB_NONE=1"
  env $E timeout 600 python bench.py --gpus $N --steps ${STEPS:-6} --warmup 2 --no-e2e --no-cpu-baseline --no-preflight --no-extra-workloads --workload $W 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   ms/step %.2f  value %.2f Grec/s  %s' % (d['ms_per_step'], d['value']/1e9, d['transport'])); print('   phases', [(k,round(v,2)) for k,v in d['phases_ms'] if v>0.15]); print('   exch', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['exchange'].items()}); print('   kern', {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
"
done
