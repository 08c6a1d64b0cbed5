#!/bin/bash
# round 2, seventh 1-GPU call: the merge probe (print fixed), and the racecheck cases that call 6 cut off --
# one test per sanitizer process, each with its own time limit and a stack dump if it sits for 4 minutes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | tr '\n' ' ')"
echo "== merge probe: 8 / 4 runs, even and 99%+1% (PROBE_SHAPE=two)"
for P in 8 4; do
  timeout 300 python tools/merge_probe.py $P 28 16 0 5
  PROBE_SHAPE=two timeout 300 python tools/merge_probe.py $P 28 16 0 5
done
timeout 300 python tools/merge_probe.py 8 27 48 2 5
PROBE_SHAPE=two timeout 300 python tools/merge_probe.py 8 27 48 2 5
echo "== racecheck, one test per process"
i=0
for T in 'test_second_sort_merge_path[12-24-desc5-2]' 'test_second_sort_merge_path[4-7-desc6-None]' 'test_all_empty_and_single_rank' 'test_host_buffers_in_chunks' 'test_multiset_hash_matches_oracle_and_sees_what_a_byte_sum_cannot'; do
  i=$((i+1))
  echo "-- $T"
  timeout 500 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --target-processes all --error-exitcode 66 --log-file gpurun_out/race_$i.log \
     python -m pytest "tests/test_gpu_parity.py::$T" -q -x -m gpu -o faulthandler_timeout=240 2>&1 | tail -40 | cut -c1-240
  echo "   exit ${PIPESTATUS[0]}; $(grep -c 'Error: Race' gpurun_out/race_$i.log) race reports; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/race_$i.log | tr '\n' ';')"
done
} 2>&1 | tee gpurun_out/call7.log
