#!/bin/bash
# round 2, tenth 1-GPU call: racecheck over the small-input slice of the suite in processes of six tests each,
# synccheck over the same slice in one process
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "# $(nvidia-smi --query-gpu=name --format=csv,noheader | tr '\n' ' ')"
export SEL='test_radix_sort_desc_matches_oracle or test_golden_vectors or test_second_sort_merge_path or test_all_empty_and_single_rank or test_record_mode_bare_8_byte_keys or test_range_compression or test_host_buffers_in_chunks or test_multiset_hash'
CHUNK=6 CHUNK_TIMEOUT=300 TOOLS="racecheck" bash tools/sanitize.sh
cp gpurun_out/sanitize_racecheck.log gpurun_out/sanitize_racecheck_chunked.log
TOOLS="synccheck" SANITIZE_TIMEOUT=400 bash tools/sanitize.sh
} 2>&1 | tee gpurun_out/call10.log
