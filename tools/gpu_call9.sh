#!/bin/bash
# round 2, ninth 1-GPU call: the GPU suite on the final code, and the one case that sits under racecheck
# (test_second_sort_merge_path[5-12-desc4-None], radix SecondSort on 5 rank threads) -- with the tiles of the
# sweeps numbered by block index (default) and by atomic ticket
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | tr '\n' ' ')"
echo "== GPU suite"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for V in "" "MPSORT_TICKET_TILES=1"; do
  echo "== racecheck, test_second_sort_merge_path[5-12-desc4-None]  $V"
  env $V MPSORT_TEST_INSTAFAIL=1 timeout 240 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --target-processes all --error-exitcode 66 --log-file gpurun_out/race9_${V%%=*}.log \
     python -X faulthandler -m pytest "tests/test_gpu_parity.py::test_second_sort_merge_path[5-12-desc4-None]" -q -x -m gpu -o faulthandler_timeout=150 2>&1 | grep -v "^$" | tail -25 | cut -c1-200
  echo "   exit ${PIPESTATUS[0]}; $(grep -c 'Error: Race' gpurun_out/race9_${V%%=*}.log) race reports; $(grep 'RACECHECK SUMMARY' gpurun_out/race9_${V%%=*}.log)"
  nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv,noheader
done
} 2>&1 | tee gpurun_out/call9.log
