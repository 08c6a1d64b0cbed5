#!/bin/bash
# A/B of the exchange transports: bash tools/exchange_modes.sh NGPUS [mode ...]
# (a mode is a quoted list of environment assignments)
N=$1; shift
if [ $# -eq 0 ]; then
  set -- "MPSORT_NO_P2P=1" "MPSORT_P2P=1 MPSORT_P2P_CE=1" "MPSORT_P2P=1 MPSORT_P2P_CE=2" "MPSORT_P2P=1 MPSORT_P2P_CE=1 MPSORT_EXCHANGE_PHASES=2"
fi
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python bench.py --gpus $N --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.2f  value %.2f Grec/s'%(d['ms_per_step'], d['value']/1e9)); print('  phases', [(k,round(v,2)) for k,v in d['phases_ms'] if v>0.2]); print('  exch', d['exchange']); print('  kern', {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
"
done
