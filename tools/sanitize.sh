#!/bin/bash
# compute-sanitizer over a small slice of the GPU suite (one gpurun call, 1 GPU, ~10 min):
#   memcheck  -- out-of-bounds / misaligned global and shared accesses, leaks of device memory
#   racecheck -- shared-memory hazards inside a CTA (the tile kernels stage through shared memory)
#   synccheck -- divergent barriers
#   initcheck -- reads of device memory nobody wrote (arena slots are reused without clearing)
# Output: gpurun_out/sanitize_<tool>.log; a summary line per tool at the end. Not part of the default
# suite: every kernel runs 10-100x slower under the tools. Written in a GPU-less session; first run is round 2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=${COMPUTE_SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
# small inputs only: record mode + hybrid needs 2^22 records (one case), everything else a few thousand
SEL=${SEL:-'test_radix_sort_desc_matches_oracle or test_radix_sort_desc_stable_on_duplicates or test_golden_vectors or test_second_sort_merge_path or test_distributed_sort_other_key_shapes or test_all_empty_and_single_rank or test_record_mode_bare_8_byte_keys or test_range_compression'}
# CHUNK=n: n tests per sanitizer process (racecheck stops making progress after some 40 tests with many rank
# threads in ONE process, profiles/r02_call6*.log, r02_call8*.log; every test passes in a process of its own)
for TOOL in ${TOOLS:-memcheck racecheck synccheck initcheck}; do
  LOG=gpurun_out/sanitize_$TOOL.log
  EXTRA=""
  [ "$TOOL" = memcheck ] && EXTRA="--leak-check full"
  [ "$TOOL" = initcheck ] && EXTRA="--track-unused-memory no"
  if [ -n "$CHUNK" ]; then
    python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$SEL" --collect-only 2>/dev/null | grep '::' > gpurun_out/sanitize_ids.txt
    split -l "$CHUNK" gpurun_out/sanitize_ids.txt gpurun_out/sanitize_chunk_
    : > $LOG; : > gpurun_out/sanitize_$TOOL.pytest.log
    NCH=0; NBAD=0
    for F in gpurun_out/sanitize_chunk_*; do
      NCH=$((NCH+1))
      MPSORT_TEST_INSTAFAIL=1 timeout ${CHUNK_TIMEOUT:-300} $CS --tool $TOOL $EXTRA --target-processes all --error-exitcode 66 --log-file $LOG.part \
          python -m pytest $(tr '\n' ' ' < $F) -q -x -m gpu >> gpurun_out/sanitize_$TOOL.pytest.log 2>&1 || { NBAD=$((NBAD+1)); echo "   chunk $F: exit $?"; }
      cat $LOG.part >> $LOG 2>/dev/null; rm -f $LOG.part $F
    done
    echo "== $TOOL in $NCH processes of <= $CHUNK tests: $NBAD failed ; $(grep -c 'passed' gpurun_out/sanitize_$TOOL.pytest.log) reported passes ; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' $LOG | sort | uniq -c | tr '\n' ';')"
    grep -h ' passed\| failed' gpurun_out/sanitize_$TOOL.pytest.log | awk '{p+=$1} END {print "   tests passed in all:", p}'
    grep -A60 '==== FAILED' gpurun_out/sanitize_$TOOL.pytest.log | cut -c1-240 | head -120
    continue
  fi
  MPSORT_TEST_INSTAFAIL=1 timeout ${SANITIZE_TIMEOUT:-1500} $CS --tool $TOOL $EXTRA --target-processes all --error-exitcode 66 --log-file $LOG \
      python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "$SEL" > gpurun_out/sanitize_$TOOL.pytest.log 2>&1
  echo "== $TOOL: exit $? ; $(grep -c 'ERROR SUMMARY' $LOG 2>/dev/null) summaries ; $(grep 'ERROR SUMMARY' $LOG 2>/dev/null | sort | uniq -c | tr '\n' ';')"
  grep -A60 '==== FAILED' gpurun_out/sanitize_$TOOL.pytest.log | cut -c1-240 | head -120
  tail -2 gpurun_out/sanitize_$TOOL.pytest.log
done | tee gpurun_out/sanitize_summary.log
