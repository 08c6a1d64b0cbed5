#!/bin/bash
# round 2, last 2-GPU call: the NCCL-process tests on the final code (own slice in place through every transport:
# DMA peer copies, peer-store kernel, pull, grouped ncclSend/ncclRecv), randomised cases as NCCL ranks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | tr '\n' ' ')"
MPSORT_TEST_CANDIDATES=1 timeout 1500 python -m pytest tests/test_zz_candidates.py tests/test_gpu_parity.py -m gpu -q -x -k "candidate or candidates or nccl or randomised" 2>&1 | tail -5
echo "== one launch: transports with the own slice in place / copied"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 tools/ab_multi.py --gpus 2 --steps 6 \
  "uniform16:-" "uniform16:MPSORT_NO_SELF_IN_PLACE=1" "mostly_sorted16:-" "mostly_sorted16:MPSORT_NO_SELF_IN_PLACE=1" 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$"
} 2>&1 | tee gpurun_out/call_n2d.log
