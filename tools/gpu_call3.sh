#!/bin/bash
# round 2, GPU call 3 (1 GPU): hash predictor + fewer host round trips, prefetch distances for every tile kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -1)"
echo "== parity of what changed (hybrid predictor/depth, fix-up, record mode, merge)"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_python_api.py -m gpu -x -q -k "hybrid or record_mode or merge or golden or full_size or multirank or radix_sort_desc or distributed or verify or sort_" 2>&1 | tail -4
echo "== record pass prefetch distance (default 148)"
for d in 0 37 74 111 148 222; do MPSORT_PREFETCH_TILES=$d timeout 150 python tools/sweep.py 28 16 0 | tail -1; done
echo "== fix-up prefetch distance"
for d in 296 592 1184 2368; do MPSORT_PREFETCH_FIXUP_TILES=$d timeout 150 python tools/sweep.py 28 16 0 | tail -1; done
echo "== bare 8-byte keys"
for d in 74 148 296; do MPSORT_PREFETCH_TILES=$d timeout 150 python tools/sweep.py 28 8 0 | tail -1; done
echo "== 48-byte particles: index pass prefetch distance"
for d in 0 74 148 296 592; do MPSORT_PREFETCH_INDEX_TILES=$d timeout 150 python tools/sweep.py 28 48 2 | tail -1; done
echo "== mostly sorted keys as rank 7 of 8 holds them: default | MPSORT_NO_HYBRID5=1"
SWEEP_AS=7,8 timeout 150 python tools/sweep.py 28 16 1 | tail -1
SWEEP_AS=7,8 MPSORT_NO_HYBRID5=1 timeout 150 python tools/sweep.py 28 16 1 | tail -1
timeout 150 python tools/sweep.py 28 16 1 | tail -1
echo "== merge alone: prefetch distance, p = 8 (generic kernel) and p = 2 (16-byte record kernel), 48-byte records"
for d in 0 74 148 296 592; do MPSORT_PREFETCH_MERGE_TILES=$d timeout 300 python tools/merge_probe.py 8 28 16 0 5 2>&1 | tail -1; done
for d in 0 148 444; do MPSORT_PREFETCH_MERGE_TILES=$d timeout 300 python tools/merge_probe.py 2 28 16 0 5 2>&1 | tail -1; done
for d in 0 148 296; do MPSORT_PREFETCH_MERGE_TILES=$d timeout 300 python tools/merge_probe.py 8 27 48 2 3 2>&1 | tail -1; done
for d in 0 148; do MPSORT_PREFETCH_MERGE_TILES=$d timeout 300 python tools/merge_probe.py 4 28 16 0 5 2>&1 | tail -1; done
} 2>&1 | tee gpurun_out/call3.log
