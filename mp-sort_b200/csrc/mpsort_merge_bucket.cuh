/*
 * mpsort_merge_bucket.cuh -- CANDIDATE (off by default, MPSORT_MERGE_BUCKET=1 selects it; not
 * yet run on a GPU): the arithmetic of a one-round tile merge by interpolation buckets.
 *
 * merge_tile_kernel orders the <= 4096 records of a tile with log2(p) merge-path rounds in
 * shared memory: 286 thread-instructions per record at p = 8, three quarters of them the rounds
 * ("merge of the received runs"). A tile, however, holds records of consecutive
 * global rank: its keys lie between two neighbouring tile-boundary sample keys [klo, khi], and
 * inside such a narrow window keys are spread almost evenly whatever the global distribution
 * is. So:   bucket(key) = (key - klo) >> sh,  sh chosen so that khi lands in the upper half of
 * NB = 8192 buckets. The map is monotone, hence the merged order is: buckets in order, and
 * inside a bucket by (key, source position). With <= 4096 records in >= 4096 used buckets a
 * bucket holds one or two records, and ranking inside a bucket is a handful of comparisons:
 *   count (shared atomics) -> exclusive scan of the counters -> scatter to bucket order ->
 *   rank inside the bucket -> sorted list of source positions.
 * Ties: the source position of a record in the receive buffer grows with (run, index in run),
 * which is the order the stable merge gives equal keys (stdlib/msort.c:78: `<=` takes from the
 * lower run), so equal keys are ranked by source position.
 * Tiles whose keys are NOT spread (duplicates: a bucket longer than CMAX; or a key outside
 * [klo, khi], which would break monotonicity) are detected after the count and take the
 * merge-path rounds instead: the bucket path can only ever be taken when it is exact.
 *
 * The functions here are plain integer arithmetic on arrays, written so that the SAME source
 * is compiled into the kernel (mpsort_kernels.cu) and into a host emulation that runs the
 * phases thread by thread (tests/native/merge_bucket_emul.cpp, driven by
 * tests/test_merge_bucket_emul.py without a GPU).
 */
#ifndef MPSORT_MERGE_BUCKET_CUH
#define MPSORT_MERGE_BUCKET_CUH

#ifdef __CUDACC__
#define MBK_HD __host__ __device__ __forceinline__
#else
#define MBK_HD static inline
#endif

namespace mbk {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr u32 NB = 8192;             /* buckets */
constexpr u32 NB_LOG2 = 13;
constexpr u32 NBP = NB + NB / 16;    /* counters, one pad word per 16: a thread scans 16 consecutive ones */
constexpr u32 CMAX = 16;             /* longest bucket that is ranked in place */

MBK_HD u32 padc(u32 b) { return b + (b >> 4); }

MBK_HD u32 bit_length(u64 x)
{
    if (x == 0) return 0;
#ifdef __CUDA_ARCH__
    return 64u - (u32) __clzll((long long) x);
#else
    return 64u - (u32) __builtin_clzll(x);
#endif
}

/* (khi - klo) >> sh < NB, and >= NB/2 whenever khi - klo >= NB/2 */
MBK_HD u32 shift_for(u64 klo, u64 khi)
{
    const u32 bits = bit_length(khi - klo);
    return bits > NB_LOG2 ? bits - NB_LOG2 : 0u;
}

/* monotone in key; < NB for klo <= key <= khi */
MBK_HD u32 bucket_of(u64 key, u64 klo, u32 sh) { return (u32) ((key - klo) >> sh); }

MBK_HD bool in_range(u64 key, u64 klo, u64 khi) { return key >= klo && key <= khi; }

/* how many members of the bucket [lo, hi) of the scattered arrays order before (key, src):
 * a smaller key, or the same key from an earlier source position */
MBK_HD u32 rank_in_bucket(const u64 * skey, const u32 * ssrc, u32 lo, u32 hi, u64 key, u32 src)
{
    u32 r = 0;
    for (u32 m = lo; m < hi; m++) {
        const u64 km = skey[m];
        const u32 sm = ssrc[m];
        r += ((km < key) || (km == key && sm < src)) ? 1u : 0u;
    }
    return r;
}

/* first position of bucket b in the scattered arrays / one past its last; `start` is the
 * exclusive scan of the counters in the padded layout, total = records in the tile */
MBK_HD u32 bucket_begin(const u32 * start, u32 b) { return start[padc(b)]; }
MBK_HD u32 bucket_end(const u32 * start, u32 b, u32 total) { return b + 1 < NB ? start[padc(b + 1)] : total; }

/* phase "rank" for the record at scattered position pos: where its source position goes in
 * the merged order */
MBK_HD u32 merged_position(const u32 * start, const u64 * skey, const u32 * ssrc, u32 pos, u32 total,
                           u64 klo, u32 sh)
{
    const u64 key = skey[pos];
    const u32 b = bucket_of(key, klo, sh);
    const u32 lo = bucket_begin(start, b), hi = bucket_end(start, b, total);
    return lo + rank_in_bucket(skey, ssrc, lo, hi, key, ssrc[pos]);
}

}  /* namespace mbk */
#endif
