#!/bin/bash
# round 2, third 2-GPU call: the parts of one exchange chained on the copy streams (no idle link behind every
# part's barrier) against the old order, 4 / 8 parts at two GPUs; NCCL parity on the final code
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | tr '\n' ' ')"
echo "== NCCL parity + fallbacks over NCCL processes + randomised NCCL ranks"
MPSORT_TEST_CANDIDATES=1 timeout 1500 python -m pytest tests/test_zz_candidates.py tests/test_gpu_parity.py -m gpu -q -x -k "candidate or candidates or nccl or randomised" 2>&1 | tail -5
echo "== one launch: chained parts (default) / old order; 4 (default at 2 GPUs) / 8 / 16 parts"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 tools/ab_multi.py --gpus 2 --steps 8 \
  "uniform16:-" "uniform16:MPSORT_NO_CHAINED_PARTS=1" "uniform16:MPSORT_EXCHANGE_PHASES=8" "uniform16:MPSORT_EXCHANGE_PHASES=8,MPSORT_NO_CHAINED_PARTS=1" \
  "uniform16:MPSORT_EXCHANGE_PHASES=16" "uniform16:MPSORT_EXCHANGE_PHASES=2" "uniform16:-" \
  "mostly_sorted16:-" "mostly_sorted16:MPSORT_NO_CHAINED_PARTS=1" "particles48:-" 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$"
echo "== bench.py as the driver runs it"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench_n2c.err | tee gpurun_out/bench_n2c.json | cut -c1-400
} 2>&1 | tee gpurun_out/call_n2c.log
