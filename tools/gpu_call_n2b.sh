#!/bin/bash
# round 2, second 2-GPU call: the rewritten peer splitter kernel (push, tagged words), sparse hint + burst copies
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | tr '\n' ' ')"
echo "== peer splitter kernel on rank threads of one GPU (opt-in tests) + NCCL parity + fallbacks over NCCL processes + randomised NCCL ranks"
MPSORT_TEST_CANDIDATES=1 timeout 1500 python -m pytest tests/test_zz_candidates.py tests/test_gpu_parity.py -m gpu -q -x -k "candidate or candidates or nccl or randomised" 2>&1 | tail -5
echo "== one launch: defaults, all-reduce splitters, mostly sorted with / without the sparse hint's choice"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546 tools/ab_multi.py --gpus 2 --steps 6 \
  "uniform16:-" "uniform16:MPSORT_NO_PEER_SPLITTER=1" "mostly_sorted16:-" "mostly_sorted16:MPSORT_EXCHANGE_PHASES=4" "mostly_sorted16:MPSORT_NO_ONE_STEP=1" "particles48:-" "particles48:MPSORT_NO_PEER_SPLITTER=1" 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$"
} 2>&1 | tee gpurun_out/call_n2b.log
