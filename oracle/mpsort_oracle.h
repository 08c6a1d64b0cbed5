/*
 * mpsort_oracle.h -- CPU restatement of MP-sort's distributed histogram sort.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load it, and only as the checker. The product (libmpsort-b200.so) has no CPU
 * path and never links or calls this.
 *
 * Parity status: PINNED. The restatement is checked (tests/test_oracle.py) against
 *   - the unmodified reference built from /root/reference with the MPI shim
 *     (oracle/_ref, recipe in oracle/Makefile), on seeded inputs incl. duplicates,
 *     empty ranks, uneven output sizes and every tuning flag, and
 *   - the golden vectors of the reference's own tests (tests/golden/: issue7,
 *     mismatched zeros, few items), which were also replayed through oracle/_ref.
 */
#ifndef MPSORT_ORACLE_H
#define MPSORT_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* key descriptor, same meaning as struct mpsort_radix_desc (include/mpsort.h);
 * raw != 0: the radix is the key bytes as they are (rsize = width*nwords, the C
 * callers' convention); raw == 0: every word widened to 8 bytes the way
 * binding.pyx:81-121 does (rsize = 8*nwords). Both orders are identical. */
struct oracle_desc { size_t offset; uint32_t width, nwords; int32_t is_signed; int32_t raw; };

#define ORACLE_DISABLE_SPARSE_ALLTOALLV (1 << 1)
#define ORACLE_DISABLE_GATHER_SORT      (1 << 3)
#define ORACLE_REQUIRE_GATHER_SORT      (1 << 4)
#define ORACLE_REQUIRE_SPARSE_ALLTOALLV (1 << 6)

/* radix_sort (radixsort.c:35-44): stable sort of one array */
void oracle_radix_sort(void * base, size_t nmemb, size_t elsize, const struct oracle_desc * d);

/* mpsort_mpi_newarray_impl (mpsort-mpi.c:161-331) for p simulated ranks.
 * bases[r]/nmemb[r]: rank r's input (left locally sorted when no gather happens,
 * like the reference); outs[r]/outnmemb[r]: rank r's output, outs[r] may equal
 * bases[r] (in place). Returns 0, or a negative code where the reference aborts.
 * Optional outputs: sendcounts[p*p] = SendCount rows of the LEADER communicator
 * mapped back to world ranks (zero rows for non-leaders), rounds = bisection
 * iterations, nleaders. */
int oracle_mpsort(int p, void ** bases, const size_t * nmemb,
        void ** outs, const size_t * outnmemb, size_t elsize,
        const struct oracle_desc * d, int options,
        int64_t * sendcounts, int * rounds, int * nleaders);

/* checksum (mpsort-mpi.c:148-159): bytes summed as signed chars into a u64 */
uint64_t oracle_checksum(const void * base, size_t nbytes);

/* the synthetic generator of the device library (oracle/synth.h) */
void oracle_generate(void * dst, size_t n, size_t elsize, int kind, uint64_t seed,
        uint64_t rank, uint64_t nranks);

#ifdef __cplusplus
}
#endif
#endif
