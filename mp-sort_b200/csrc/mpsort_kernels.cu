/*
 * mpsort_kernels.cu -- hand-written sm_100a kernels of mpsort-b200.
 *
 * All kernels here are HBM-bound integer/byte work (no tensor cores: no stage of a
 * sort is a dense contraction). Design rules applied throughout:
 *   - coalesced warp-striped loads, shared-memory staging so that every global
 *     store is a run of consecutive addresses,
 *   - grids sized as multiples of the SM count for the streaming kernels,
 *   - no global atomics on the hot path other than one ticket per tile and the
 *     decoupled look-back status words.
 *
 * One translation unit; the sections live in kernels/*.cuh and are included below in
 * dependency order (common helpers first).
 *
 * Kernel <-> reference map (paths relative to MP-sort v0.1.19):
 *   extract_kernel, rec_hist_kernel   radix() callbacks               binding.pyx:81-121, bench-mpi.c:13-15   [extract_hist.cuh]
 *   onesweep_kernel, onesweep_rec_kernel, fixup_rec_kernel
 *                                     mpsort_qsort_r / msort_with_tmp stdlib/msort.c:52-174,177-314
 *   gather_records_kernel             record moves of the merge sort  stdlib/msort.c:153-173,270-294
 *   splitter_*_kernel                 _histogram/_bsearch_last_lt/le  internal-parallel.h:8-126
 *   merge_*_kernel                    the second radix_sort           mpsort-mpi.c:597
 *   p2p_copy_kernel                   MPIU_Alltoallv                  mp-mpiu.c:69-236
 *   checksum_kernel                   checksum()                      mpsort-mpi.c:148-159
 * Candidates, off by default (each names its switch where it is defined):
 *   splitter_descent_peer_kernel      the bisection loop + Allreduce  mpsort-mpi.c:385-437
 *   p2p_gather_kernel                 record moves + MPIU_Alltoallv   stdlib/msort.c:153-173, mp-mpiu.c:69-236
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "kernels/common.cuh"
#include "kernels/extract_hist.cuh"
#include "kernels/onesweep.cuh"
#include "kernels/onesweep_rec.cuh"
#include "kernels/hybrid_fixup.cuh"
#include "kernels/gather.cuh"
#include "kernels/splitter.cuh"
#include "kernels/merge.cuh"
#include "kernels/p2p_exchange.cuh"
#include "kernels/checksum.cuh"
#include "kernels/support.cuh"
