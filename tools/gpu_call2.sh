#!/bin/bash
# round 2, GPU call 2 (1 GPU): static tiles by default, L2 prefetch distances, fix-up v2, merge with ranked samples
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -1)"
echo "== parity of what changed (hybrid fix-up, record mode, merge, golden vectors)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hybrid or record_mode or merge or golden or full_size or multirank or radix_sort_desc or distributed" 2>&1 | tail -4
echo "== record pass: default (static tiles) | tickets | L2 prefetch of tile + d"
timeout 150 python tools/sweep.py 28 16 0 | tail -1
MPSORT_TICKET_TILES=1 timeout 150 python tools/sweep.py 28 16 0 | tail -1
for d in 148 296 444 592 888 1776; do MPSORT_PREFETCH_TILES=$d timeout 150 python tools/sweep.py 28 16 0 | tail -1; done
echo "== bare 8-byte keys, 48-byte particles (index mode)"
timeout 150 python tools/sweep.py 28 8 0 | tail -1
MPSORT_PREFETCH_TILES=444 timeout 150 python tools/sweep.py 28 8 0 | tail -1
timeout 150 python tools/sweep.py 28 48 2 | tail -1
MPSORT_TICKET_TILES=1 timeout 150 python tools/sweep.py 28 48 2 | tail -1
echo "== mostly sorted keys as rank 7 of 8 holds them: default | MPSORT_HYBRID_DEPTH5=1"
SWEEP_AS=7,8 timeout 150 python tools/sweep.py 28 16 1 | tail -1
SWEEP_AS=7,8 MPSORT_HYBRID_DEPTH5=1 timeout 150 python tools/sweep.py 28 16 1 | tail -1
echo "== merge alone: ranked samples (3 launches) | round-1 sample sort"
for P in 8 4 2; do
  timeout 300 python tools/merge_probe.py $P 28 16 0 5 2>&1 | tail -1
  MPSORT_LIB=$PWD/mp-sort_b200/variants/libmpsort-b200.samplesort.so timeout 300 python tools/merge_probe.py $P 28 16 0 5 2>&1 | tail -1
done
timeout 300 python tools/merge_probe.py 8 25 16 0 5 2>&1 | tail -1
MPSORT_LIB=$PWD/mp-sort_b200/variants/libmpsort-b200.samplesort.so timeout 300 python tools/merge_probe.py 8 25 16 0 5 2>&1 | tail -1
timeout 300 python tools/merge_probe.py 8 27 48 2 3 2>&1 | tail -1
echo "== randomised cases against the oracle on the real library (opt-in test, rank threads)"
MPSORT_TEST_CANDIDATES=1 timeout 900 python -m pytest tests/test_zz_candidates.py -m gpu -q -k "randomised and rank-threads" 2>&1 | tail -3
} 2>&1 | tee gpurun_out/call2.log
