#!/bin/bash
# round 2, eighth 1-GPU call: own slice merged in place (kernels read one run from another base) -- the GPU suite,
# the merge probe, and the racecheck sequence that call 6 lost (failures now written out at once)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | tr '\n' ' ')"
echo "== GPU suite"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== merge probe"
timeout 300 python tools/merge_probe.py 8 28 16 0 5
timeout 300 python tools/merge_probe.py 2 28 16 0 5
echo "== 4 rank threads on one GPU, 2^24 records each: own slice in place / copied"
timeout 300 python tools/group_probe.py 4 24 16 1 2>&1 | tail -3
MPSORT_NO_SELF_IN_PLACE=1 timeout 300 python tools/group_probe.py 4 24 16 1 2>&1 | tail -3
echo "== racecheck over the golden vectors and the merge path, in one process as in call 6"
SEL='test_golden_vectors or test_second_sort_merge_path or test_all_empty_and_single_rank' TOOLS="racecheck" SANITIZE_TIMEOUT=700 bash tools/sanitize.sh
echo "== memcheck over the merge path (own slice from another base)"
SEL='test_second_sort_merge_path or test_randomised_cases_against_the_oracle' TOOLS="memcheck" SANITIZE_TIMEOUT=400 bash tools/sanitize.sh
} 2>&1 | tee gpurun_out/call8.log
