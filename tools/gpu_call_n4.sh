#!/bin/bash
# round 2, one 4-GPU call: exchange parts at four GPUs (4 is the default below eight ranks), gate modes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c | tr '\n' ' ')"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 tools/ab_multi.py --gpus 4 --steps 8 \
  "uniform16:-" "uniform16:MPSORT_EXCHANGE_PHASES=8" "uniform16:MPSORT_EXCHANGE_PHASES=6" "uniform16:MPSORT_CHAINED_PARTS=1" "uniform16:MPSORT_EXCHANGE_PHASES=8,MPSORT_CHAINED_PARTS=1" "uniform16:-" 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$"
} 2>&1 | tee gpurun_out/call_n4.log
