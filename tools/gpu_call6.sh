#!/bin/bash
# round 2, sixth 1-GPU call: the merge rounds over non-empty runs only (probe: even runs vs. the two runs of a
# mostly sorted exchange), the whole GPU suite on the final kernels, compute-sanitizer over the kernels new this round
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | tr '\n' ' ')"
echo "== merge probe: 8 / 4 runs, even and 99%+1% (PROBE_SHAPE=two)"
for P in 8 4; do
  timeout 300 python tools/merge_probe.py $P 28 16 0 5
  PROBE_SHAPE=two timeout 300 python tools/merge_probe.py $P 28 16 0 5
done
PROBE_SHAPE=two timeout 300 python tools/merge_probe.py 8 27 48 2 5
timeout 300 python tools/merge_probe.py 8 27 48 2 5
echo "== GPU suite"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== compute-sanitizer (memcheck, racecheck) over the kernels new this round"
SEL='test_radix_sort_desc_matches_oracle or test_golden_vectors or test_second_sort_merge_path or test_all_empty_and_single_rank or test_record_mode_bare_8_byte_keys or test_range_compression or test_host_buffers_in_chunks or test_multiset_hash' \
  TOOLS="memcheck racecheck" SANITIZE_TIMEOUT=900 bash tools/sanitize.sh
} 2>&1 | tee gpurun_out/call6.log
