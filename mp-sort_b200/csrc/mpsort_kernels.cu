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
 * Kernel <-> reference map (paths relative to MP-sort v0.1.19):
 *   extract_kernel, rec_hist_kernel   radix() callbacks               binding.pyx:81-121, bench-mpi.c:13-15
 *   onesweep_kernel, onesweep_rec_kernel, fixup_rec_kernel
 *                                     mpsort_qsort_r / msort_with_tmp stdlib/msort.c:52-174,177-314
 *   gather_records_kernel             record moves of the merge sort  stdlib/msort.c:153-173,270-294
 *   splitter_*_kernel                 _histogram/_bsearch_last_lt/le  internal-parallel.h:8-126
 *   merge_*_kernel                    the second radix_sort           mpsort-mpi.c:597
 *   p2p_copy_kernel                   MPIU_Alltoallv                  mp-mpiu.c:69-236
 *   checksum_kernel                   checksum()                      mpsort-mpi.c:148-159
 * Candidates, off by default (each names its switch where it is defined):
 *   merge_tile_bucket_kernel          the second radix_sort           mpsort-mpi.c:597
 *   splitter_descent_peer_kernel      the bisection loop + Allreduce  mpsort-mpi.c:385-437
 *   p2p_gather_kernel                 record moves + MPIU_Alltoallv   stdlib/msort.c:153-173, mp-mpiu.c:69-236
 *   onesweep_rec_persist_kernel       mpsort_qsort_r (compile-time)   stdlib/msort.c:177-314
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "mpsort_kernels.h"
#include "mpsort_merge_bucket.cuh"

typedef unsigned long long u64;
typedef unsigned int u32;

#define FULL_MASK 0xffffffffu

static int g_num_sms = 0;
static int num_sms()
{
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

/* every kernel launch of this library passes through here: counted for bench.py's
 * "gpu_launches" claim */
static unsigned long long g_launches = 0;
#define CUDA_LAUNCH_CHECK() do { __atomic_fetch_add(&g_launches, 1ULL, __ATOMIC_RELAXED); \
    cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int) e__; } while (0)

extern "C" uint64_t mpsk_launch_count(int reset)
{
    const unsigned long long v = __atomic_load_n(&g_launches, __ATOMIC_RELAXED);
    if (reset) __atomic_store_n(&g_launches, 0ULL, __ATOMIC_RELAXED);
    return (uint64_t) v;
}

__device__ __forceinline__ u32 lanemask_lt()
{
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

/* ========================================================================= */
/* K1: key extraction + 8 digit histograms                                   */
/* ========================================================================= */

struct KeyDesc {
    size_t elsize;
    size_t offset;
    u32 width;
    u32 nwords;
    int is_signed;
    u32 g;          /* which packed 64-bit word to produce */
    u64 sub;        /* subtracted from the packed word (range compression); 0 otherwise */
};

/* little-endian load of `width` bytes, alignment-safe */
__device__ __forceinline__ u64 load_narrow(const unsigned char * p, u32 width)
{
    switch (width) {
        case 8:
            if ((((uintptr_t) p) & 7) == 0) return *(const u64 *) p;
            break;
        case 4:
            if ((((uintptr_t) p) & 3) == 0) return *(const u32 *) p;
            break;
        case 2:
            if ((((uintptr_t) p) & 1) == 0) return *(const unsigned short *) p;
            break;
        default:
            return *p;
    }
    u64 v = 0;
    for (u32 b = 0; b < width; b++) v |= ((u64) p[b]) << (8 * b);
    return v;
}

/* Packed 64-bit word g of the key of one record: key bytes [8g, 8g+8) of the
 * little-endian byte string formed by the (sign-flipped) key words. */
__device__ __forceinline__ u64 pack_key_word(const unsigned char * rec, const KeyDesc & d)
{
    const u32 per = 8 / d.width;
    const u32 first = d.g * per;
    u64 out = 0;
#pragma unroll 1
    for (u32 k = 0; k < per; k++) {
        const u32 wi = first + k;
        if (wi >= d.nwords) break;
        u64 v = load_narrow(rec + d.offset + (size_t) wi * d.width, d.width);
        if (d.is_signed) v ^= 1ULL << (8 * d.width - 1);
        out |= v << (8 * d.width * k);
    }
    return out;
}

/* fast path of the benchmark configs: one aligned 8-byte word */
__device__ __forceinline__ u64 load_key_fast8(const unsigned char * rec, size_t offset, u64 flip)
{
    return (*(const u64 *) (rec + offset)) ^ flip;
}

/* Four keys per thread and iteration. Digits that are equal over all 128 keys of the
 * warp's batch (small ids, zero high bytes, sorted input) would serialise same-address
 * shared atomics: one OR-reduction of the pairwise differences per batch finds them,
 * lane 0 adds 128 for those, everyone adds 1 per key for the rest. */
#define EXTRACT_BATCH 4
template <bool FAST8, bool MINMAX, bool INPLACE>
__global__ void __launch_bounds__(512)
extract_kernel(const unsigned char * __restrict__ base, size_t n, KeyDesc d,
               u64 * kout, u32 * __restrict__ hist, u64 * __restrict__ minmax)
{
    __shared__ u32 sh[8 * 256];
    for (u32 t = threadIdx.x; t < 8 * 256; t += blockDim.x) sh[t] = 0;
    __syncthreads();

    const u64 flip = (d.is_signed ? (1ULL << 63) : 0ULL);
    const size_t per_block = (size_t) blockDim.x * EXTRACT_BATCH;
    const size_t nblocks_total = (n + per_block - 1) / per_block;
    const bool lane0 = (threadIdx.x & 31) == 0;
    u64 kmin = ~0ULL, kmax = 0ULL;
    for (size_t blk = blockIdx.x; blk < nblocks_total; blk += gridDim.x) {
        const size_t i0 = blk * per_block + threadIdx.x;
        u64 k[EXTRACT_BATCH];
        bool valid[EXTRACT_BATCH];
#pragma unroll
        for (int j = 0; j < EXTRACT_BATCH; j++) {
            const size_t i = i0 + (size_t) j * blockDim.x;
            valid[j] = i < n;
            k[j] = 0;
            if (valid[j]) {
                const unsigned char * rec = base + i * d.elsize;
                if (INPLACE) k[j] = kout[i];              /* rebase pass: bare u64 keys, rewritten in place */
                else if (FAST8) k[j] = load_key_fast8(rec, d.offset, flip);
                else k[j] = pack_key_word(rec, d);
                k[j] -= d.sub;
                if (kout) kout[i] = k[j];
                if (MINMAX) {
                    kmin = k[j] < kmin ? k[j] : kmin;
                    kmax = k[j] > kmax ? k[j] : kmax;
                }
            }
        }
        /* the last lane's last key is the first to fall off the end */
        const bool full = __all_sync(FULL_MASK, valid[EXTRACT_BATCH - 1]);
        u32 same = 0;
        if (full) {
            const u64 k0 = __shfl_sync(FULL_MASK, k[0], 0);
            u64 diff = 0;
#pragma unroll
            for (int j = 0; j < EXTRACT_BATCH; j++) diff |= k[j] ^ k0;
            const u32 dlo = __reduce_or_sync(FULL_MASK, (u32) diff);
            const u32 dhi = __reduce_or_sync(FULL_MASK, (u32) (diff >> 32));
#pragma unroll
            for (int dd = 0; dd < 4; dd++) {
                if (((dlo >> (8 * dd)) & 255u) == 0) same |= 1u << dd;
                if (((dhi >> (8 * dd)) & 255u) == 0) same |= 1u << (dd + 4);
            }
        }
#pragma unroll
        for (int dd = 0; dd < 8; dd++) {
            if (same & (1u << dd)) {
                if (lane0) atomicAdd(&sh[dd * 256 + ((u32) (k[0] >> (8 * dd)) & 255u)], 32u * EXTRACT_BATCH);
            } else {
#pragma unroll
                for (int j = 0; j < EXTRACT_BATCH; j++)
                    if (valid[j]) atomicAdd(&sh[dd * 256 + ((u32) (k[j] >> (8 * dd)) & 255u)], 1u);
            }
        }
    }
    __syncthreads();
    for (u32 t = threadIdx.x; t < 8 * 256; t += blockDim.x) {
        const u32 c = sh[t];
        if (c) atomicAdd(&hist[t], c);
    }
    if (MINMAX) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const u64 a = __shfl_xor_sync(FULL_MASK, kmin, o), b = __shfl_xor_sync(FULL_MASK, kmax, o);
            kmin = a < kmin ? a : kmin;
            kmax = b > kmax ? b : kmax;
        }
        if (lane0) { atomicMin(&minmax[0], kmin); atomicMax(&minmax[1], kmax); }
    }
}

extern "C" int mpsk_extract_keys(const void * base, size_t n, size_t elsize,
        size_t offset, uint32_t width, uint32_t nwords, int is_signed,
        uint32_t g, uint64_t sub, uint64_t * kout, uint32_t * hist, uint64_t * minmax, mpsk_stream_t stream)
{
    if (n == 0) return 0;
    KeyDesc d;
    d.elsize = elsize; d.offset = offset; d.width = width; d.nwords = nwords;
    d.is_signed = is_signed; d.g = g; d.sub = sub;
    const int threads = 512;
    size_t blocks = (n + (size_t) threads * EXTRACT_BATCH - 1) / ((size_t) threads * EXTRACT_BATCH);
    const size_t maxb = (size_t) num_sms() * 8;
    if (blocks > maxb) blocks = maxb;
    const bool fast8 = (width == 8) && (nwords >= 1) && (elsize % 8 == 0)
                       && (offset % 8 == 0) && ((((uintptr_t) base) & 7) == 0);
    const unsigned grid = (unsigned) blocks;
    cudaStream_t st = (cudaStream_t) stream;
    const unsigned char * pb = (const unsigned char *) base;
    if (base == (const void *) kout && elsize == 8) {
        /* in-place rebase of bare u64 keys */
        extract_kernel<true, false, true><<<grid, threads, 0, st>>>(pb, n, d, (u64 *) kout, hist, (u64 *) minmax);
    } else if (fast8) {
        /* word g of an 8-byte-word key is simply word g */
        d.offset = offset + (size_t) g * 8;
        if (minmax) extract_kernel<true, true, false><<<grid, threads, 0, st>>>(pb, n, d, (u64 *) kout, hist, (u64 *) minmax);
        else extract_kernel<true, false, false><<<grid, threads, 0, st>>>(pb, n, d, (u64 *) kout, hist, (u64 *) minmax);
    } else {
        if (minmax) extract_kernel<false, true, false><<<grid, threads, 0, st>>>(pb, n, d, (u64 *) kout, hist, (u64 *) minmax);
        else extract_kernel<false, false, false><<<grid, threads, 0, st>>>(pb, n, d, (u64 *) kout, hist, (u64 *) minmax);
    }
    CUDA_LAUNCH_CHECK();
    return 0;
}

/*
 * Record mode (keys sit in place inside 8- or 16-byte records): digit histograms of NH
 * consecutive digits d0 .. d0+NH-1 plus the OR of (key ^ key[0]) over all keys. The
 * OR tells exactly which key bytes vary; the hybrid sort only ever needs the counts of
 * the four most significant digits, and counting four digits instead of eight takes
 * the kernel from shared-atomic-bound (1.0 ms per 2^28 keys, ncu: LSU wavefronts 88 %)
 * to the HBM read time. Same batch trick as extract_kernel for digits that are equal
 * over a warp's 128 keys.
 */
template <int NH>
__global__ void __launch_bounds__(512)
rec_hist_kernel(const u64 * __restrict__ words, u32 W, u32 koff, size_t n, u64 flip, u32 d0,
                u32 * __restrict__ hist, unsigned long long * __restrict__ diff)
{
    __shared__ u32 sh[NH * 256];
    for (u32 t = threadIdx.x; t < NH * 256; t += blockDim.x) sh[t] = 0;
    __syncthreads();
    const u64 k0 = words[koff] ^ flip;
    const size_t per_block = (size_t) blockDim.x * EXTRACT_BATCH;
    const size_t nblocks_total = (n + per_block - 1) / per_block;
    const bool lane0 = (threadIdx.x & 31) == 0;
    const u32 sh0 = 8 * d0;
    u64 acc = 0;
    for (size_t blk = blockIdx.x; blk < nblocks_total; blk += gridDim.x) {
        const size_t i0 = blk * per_block + threadIdx.x;
        u64 k[EXTRACT_BATCH];
        bool valid[EXTRACT_BATCH];
#pragma unroll
        for (int j = 0; j < EXTRACT_BATCH; j++) {
            const size_t i = i0 + (size_t) j * blockDim.x;
            valid[j] = i < n;
            k[j] = valid[j] ? (words[(size_t) W * i + koff] ^ flip) : k0;
        }
        u64 d = 0;
#pragma unroll
        for (int j = 0; j < EXTRACT_BATCH; j++) d |= k[j] ^ k0;
        acc |= d;
        const bool full = __all_sync(FULL_MASK, valid[EXTRACT_BATCH - 1]);
        u32 same = 0;
        if (full) {
            /* digits equal over the whole batch: compare with the warp's first key */
            const u64 kw0 = __shfl_sync(FULL_MASK, k[0], 0);
            u64 dd = 0;
#pragma unroll
            for (int j = 0; j < EXTRACT_BATCH; j++) dd |= k[j] ^ kw0;
            dd >>= sh0;
            const u32 dlo = __reduce_or_sync(FULL_MASK, (u32) dd);
            const u32 dhi = NH > 4 ? __reduce_or_sync(FULL_MASK, (u32) (dd >> 32)) : 0u;
#pragma unroll
            for (int q = 0; q < NH; q++) {
                const u32 byte = q < 4 ? ((dlo >> (8 * q)) & 255u) : ((dhi >> (8 * (q - 4))) & 255u);
                if (byte == 0) same |= 1u << q;
            }
        }
#pragma unroll
        for (int q = 0; q < NH; q++) {
            if (same & (1u << q)) {
                if (lane0) atomicAdd(&sh[q * 256 + ((u32) (k[0] >> (sh0 + 8 * q)) & 255u)], 32u * EXTRACT_BATCH);
            } else {
#pragma unroll
                for (int j = 0; j < EXTRACT_BATCH; j++)
                    if (valid[j]) atomicAdd(&sh[q * 256 + ((u32) (k[j] >> (sh0 + 8 * q)) & 255u)], 1u);
            }
        }
    }
    __syncthreads();
    for (u32 t = threadIdx.x; t < NH * 256; t += blockDim.x) {
        const u32 c = sh[t];
        if (c) atomicAdd(&hist[d0 * 256 + t], c);
    }
    if (diff) {
        const u32 lo = __reduce_or_sync(FULL_MASK, (u32) acc), hi = __reduce_or_sync(FULL_MASK, (u32) (acc >> 32));
        if (lane0 && (lo | hi)) atomicOr(diff, ((unsigned long long) hi << 32) | lo);
    }
}

/* OR of (key ^ key[0]) over s evenly spaced records: a cheap preview of which key bytes vary */
__global__ void __launch_bounds__(256)
rec_sample_diff_kernel(const u64 * __restrict__ words, u32 W, u32 koff, size_t n, u32 s, unsigned long long * __restrict__ diff)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    u64 acc = 0;
    if (i < s) {
        const size_t pos = (size_t) (((unsigned __int128) i * n) / s);
        acc = words[(size_t) W * pos + koff] ^ words[koff];
    }
    const u32 lo = __reduce_or_sync(FULL_MASK, (u32) acc), hi = __reduce_or_sync(FULL_MASK, (u32) (acc >> 32));
    if ((threadIdx.x & 31) == 0 && (lo | hi)) atomicOr(diff, ((unsigned long long) hi << 32) | lo);
}

extern "C" int mpsk_rec_histograms(const void * recs, size_t n, size_t elsize, int key_in_high, uint64_t flip,
        uint32_t d0, uint32_t nh, uint32_t * hist, uint64_t * diff, mpsk_stream_t stream)
{
    if (n == 0) return 0;
    if ((elsize != 8 && elsize != 16) || (nh != 4 && nh != 8) || d0 + nh > 8) return (int) cudaErrorInvalidValue;
    const int threads = 512;
    size_t blocks = (n + (size_t) threads * EXTRACT_BATCH - 1) / ((size_t) threads * EXTRACT_BATCH);
    const size_t maxb = (size_t) num_sms() * 8;
    if (blocks > maxb) blocks = maxb;
    const u32 W = (u32) (elsize / 8), koff = (key_in_high && elsize == 16) ? 1u : 0u;
    if (nh == 4)
        rec_hist_kernel<4><<<(unsigned) blocks, threads, 0, (cudaStream_t) stream>>>(
            (const u64 *) recs, W, koff, n, (u64) flip, d0, hist, (unsigned long long *) diff);
    else
        rec_hist_kernel<8><<<(unsigned) blocks, threads, 0, (cudaStream_t) stream>>>(
            (const u64 *) recs, W, koff, n, (u64) flip, d0, hist, (unsigned long long *) diff);
    CUDA_LAUNCH_CHECK();
    return 0;
}

extern "C" int mpsk_rec_sample_diff(const void * recs, size_t n, size_t elsize, int key_in_high, uint32_t s,
        uint64_t * diff, mpsk_stream_t stream)
{
    if (n == 0 || s == 0) return 0;
    if (elsize != 8 && elsize != 16) return (int) cudaErrorInvalidValue;
    rec_sample_diff_kernel<<<(s + 255) / 256, 256, 0, (cudaStream_t) stream>>>(
        (const u64 *) recs, (u32) (elsize / 8), (key_in_high && elsize == 16) ? 1u : 0u, n, s, (unsigned long long *) diff);
    CUDA_LAUNCH_CHECK();
    return 0;
}

/* exclusive scan of nhist 256-bin histograms, one warp-synchronous block each */
__global__ void __launch_bounds__(256)
scan_hist_kernel(const u32 * __restrict__ hist, u32 * __restrict__ bins)
{
    __shared__ u32 wsum[8];
    const u32 t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const u32 c = hist[blockIdx.x * 256 + t];
    u32 incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 y = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    u32 add = 0;
    for (u32 w = 0; w < warp; w++) add += wsum[w];
    bins[blockIdx.x * 256 + t] = incl - c + add;
}

extern "C" int mpsk_scan_histograms(const uint32_t * hist, uint32_t * bins, int nhist, mpsk_stream_t stream)
{
    if (nhist <= 0) return 0;
    scan_hist_kernel<<<nhist, 256, 0, (cudaStream_t) stream>>>(hist, bins);
    CUDA_LAUNCH_CHECK();
    return 0;
}

/* ========================================================================= */
/* K2: onesweep pass                                                         */
/* ========================================================================= */

#ifndef MPSK_SWEEP_THREADS
#define MPSK_SWEEP_THREADS 384
#endif
#ifndef MPSK_SWEEP_IPT
#define MPSK_SWEEP_IPT 16
#endif
#ifndef MPSK_SWEEP_MINBLOCKS
#define MPSK_SWEEP_MINBLOCKS 2
#endif
#ifndef MPSK_USE_MATCH
#define MPSK_USE_MATCH 0
#endif

constexpr u32 LB_PART = 1u << 30;
constexpr u32 LB_INCL = 2u << 30;
constexpr u32 LB_MASK = (1u << 30) - 1u;

__device__ __forceinline__ u32 ld_relaxed_u32(const u32 * p)
{
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(u32 * p, u32 v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

/* lanes of the warp whose digit equals mine.
 * __match_any_sync (MATCH.ANY) costs ~2x more MIO time than eight ballots on B200
 * (profiles/r01_sweep1_match_vs_ballot.log), so the default splits bit by bit:
 * per bit one predicate, one VOTE and one predicated AND. */
template <int BIT>
__device__ __forceinline__ u32 match_bit(u32 peers, u32 digit)
{
    asm("{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 t, v;\n\t"
        "and.b32 t, %1, %2;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 v, p, 0xffffffff;\n\t"
        "@!p not.b32 v, v;\n\t"
        "and.b32 %0, %0, v;\n\t"
        "}" : "+r"(peers) : "r"(digit), "n"(1u << BIT));
    return peers;
}

__device__ __forceinline__ u32 match_digit(u32 digit)
{
#if MPSK_USE_MATCH
    return __match_any_sync(FULL_MASK, digit);
#else
    u32 peers = FULL_MASK;
    peers = match_bit<0>(peers, digit);
    peers = match_bit<1>(peers, digit);
    peers = match_bit<2>(peers, digit);
    peers = match_bit<3>(peers, digit);
    peers = match_bit<4>(peers, digit);
    peers = match_bit<5>(peers, digit);
    peers = match_bit<6>(peers, digit);
    peers = match_bit<7>(peers, digit);
    return peers;
#endif
}


/*
 * Decoupled look-back of one (tile, digit): exclusive count of the digit over all
 * earlier tiles, two levels deep.
 *
 * With a flat look-back every in-flight predecessor only has a PARTIAL count until
 * its own walk ends, and with ~450 small tiles resident the walk was ~160 entries
 * deep: ncu showed 25 % of all instructions of a pass in this loop
 * (profiles/r01_ncu_rec16_flat_lookback.txt). Tiles are therefore grouped in blocks
 * of LB_BLOCK consecutive tiles. Every tile also adds its count to its block's total
 * with ONE atomic that carries an arrival counter in the top bits
 * ({arrivals:6, count:26}), so a complete block is a single self-describing word.
 * A walk covers at most LB_BLOCK-1 tiles of its own block and then whole blocks.
 * All waits are on tiles with smaller tickets, which are running or done.
 */
#ifndef MPSK_LB_DEPTH
#define MPSK_LB_DEPTH 4
#endif
#ifndef MPSK_LB_BLOCK
#define MPSK_LB_BLOCK 32
#endif
constexpr int LB_BLOCK = MPSK_LB_BLOCK;
constexpr u32 LB_TOTAL_SHIFT = 26;
constexpr u32 LB_TOTAL_MASK = (1u << LB_TOTAL_SHIFT) - 1u;

struct LookbackBufs { u32 * tiles; u32 * blktotal; u32 * blkincl; };

__device__ __forceinline__ void lookback_publish_partial(const LookbackBufs & lb, u32 tile, u32 digit, u32 count)
{
    st_relaxed_u32(&lb.tiles[(size_t) tile * 256 + digit], (tile == 0 ? LB_INCL : LB_PART) | count);
    atomicAdd(&lb.blktotal[(size_t) (tile / LB_BLOCK) * 256 + digit], (1u << LB_TOTAL_SHIFT) | count);
}

/* walk the tile entries t, t-1, ..., t_first (MPSK_LB_DEPTH polled per round trip);
 * true when an INCLUSIVE entry ended the walk */
__device__ __forceinline__ bool lookback_walk(const u32 * tiles, int t, const int t_first, const u32 digit, u32 & acc)
{
    while (t >= t_first) {
        u32 s[MPSK_LB_DEPTH];
#pragma unroll
        for (int k = 0; k < MPSK_LB_DEPTH; k++)
            s[k] = (t - k >= t_first) ? ld_relaxed_u32(&tiles[(size_t) (t - k) * 256 + digit]) : 0u;
        int used = 0;
#pragma unroll
        for (int k = 0; k < MPSK_LB_DEPTH; k++) {
            if (used == k && t - k >= t_first) {
                if (s[k] & LB_INCL) { acc += s[k] & LB_MASK; return true; }
                if (s[k] & LB_PART) { acc += s[k] & LB_MASK; used++; }
            }
        }
        t -= used;                     /* entries not yet published are polled again */
    }
    return false;
}

__device__ __forceinline__ u32 lookback_exclusive(const LookbackBufs & lb, u32 tile, u32 digit)
{
    const u32 b = tile / LB_BLOCK;
    u32 excl = 0;
    /* 1: the earlier tiles of my own block */
    if (lookback_walk(lb.tiles, (int) tile - 1, (int) (b * LB_BLOCK), digit, excl)) return excl;
    if (b == 0) return excl;
    /* 2: whole blocks, newest first */
    u32 e2 = 0;
    int bb = (int) b - 1;
    for (;;) {
        const u32 wi = ld_relaxed_u32(&lb.blkincl[(size_t) bb * 256 + digit]);
        const u32 wt = ld_relaxed_u32(&lb.blktotal[(size_t) bb * 256 + digit]);
        if (wi & LB_INCL) { e2 += wi & LB_MASK; break; }
        if ((wt >> LB_TOTAL_SHIFT) == (u32) LB_BLOCK) {
            e2 += wt & LB_TOTAL_MASK;
        } else {
            /* some tile of that block has not ranked yet: take its tiles one by one */
            if (lookback_walk(lb.tiles, (bb + 1) * LB_BLOCK - 1, bb * LB_BLOCK, digit, e2)) break;
        }
        if (bb == 0) break;
        bb--;
    }
    /* e2 is the inclusive prefix through block b-1: later walks stop here */
    st_relaxed_u32(&lb.blkincl[(size_t) (b - 1) * 256 + digit], LB_INCL | e2);
    return excl + e2;
}

__host__ __device__ __forceinline__ size_t lookback_words(size_t ntiles)
{
    const size_t nblk = (ntiles + LB_BLOCK - 1) / LB_BLOCK;
    return 64 + ntiles * 256 + 2 * nblk * 256;
}

template <int THREADS, int IPT>
struct SweepCfg {
    static constexpr int TILE = THREADS * IPT;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int VAL_BYTES = (TILE * 4 > WARPS * 256 * 4) ? TILE * 4 : WARPS * 256 * 4;
    static constexpr int SMEM = TILE * 8 + VAL_BYTES + 256 * 4 * 2 + 64;
};

/*
 * One CTA sorts one tile of TILE (key,value) pairs by the 8-bit digit at `shift`
 * and appends every digit's run to that digit's global output region. The global
 * position of a tile's run is  bins[d] (all smaller digits, whole array)
 *                            + sum over earlier tiles of their count of digit d,
 * the second term found with a decoupled look-back over per-(tile,digit) status
 * words {flag:2, count:30}. Tiles take a ticket so that every predecessor of a
 * running tile has itself started (forward progress of the spin).
 *
 * Stability: tile order = ticket order = input order; inside a tile items are
 * ranked in (warp, round j, lane) order, which is exactly the order they were
 * loaded in (position = warp*IPT*32 + j*32 + lane).
 */
template <int THREADS, int IPT, bool IOTA>
__global__ void __launch_bounds__(THREADS, MPSK_SWEEP_MINBLOCKS)
onesweep_kernel(const u64 * __restrict__ kin, const u32 * __restrict__ vin,
                u64 * __restrict__ kout, u32 * __restrict__ vout,
                u32 n, u32 shift, const u32 * __restrict__ bins,
                LookbackBufs lb, u32 * ticket)
{
    typedef SweepCfg<THREADS, IPT> Cfg;
    constexpr int TILE = Cfg::TILE;
    constexpr int WARPS = Cfg::WARPS;
    static_assert(THREADS >= 256, "one thread per digit needs >= 256 threads");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 * s_keys = (u64 *) smem_raw;
    u32 * s_vals = (u32 *) (smem_raw + TILE * 8);
    u32 * s_whist = s_vals;   /* [WARPS][256]; dead before values are staged */
    u32 * s_local = (u32 *) (smem_raw + TILE * 8 + Cfg::VAL_BYTES);
    u32 * s_gofs = s_local + 256;
    u32 * s_misc = s_gofs + 256;  /* [0] tile id, [1..8] digit-scan warp totals */

    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31u;
    const u32 warp = tid >> 5;

    if (tid == 0) s_misc[0] = atomicAdd(ticket, 1u);
    for (u32 i = tid; i < WARPS * 256; i += THREADS) s_whist[i] = 0;
    __syncthreads();

    const u32 tile = s_misc[0];
    const u32 tile_base = tile * (u32) TILE;
    const u32 remaining = n - tile_base;
    const u32 valid = remaining < (u32) TILE ? remaining : (u32) TILE;
    const u32 wbase = tile_base + warp * (IPT * 32) + lane;

    /* ---- load keys, warp-striped: each load instruction covers 256 contiguous bytes */
    u64 key[IPT];
    if (valid == (u32) TILE) {
#pragma unroll
        for (int j = 0; j < IPT; j++) key[j] = kin[wbase + j * 32];
    } else {
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            const u32 pos = wbase + j * 32;
            key[j] = pos < n ? kin[pos] : ~0ULL;   /* padding ranks last in bin 255 */
        }
    }

    /* ---- rank inside (warp, digit): match peers, leader bumps the warp counter */
    u32 rank[IPT];
    u32 * my_hist = s_whist + warp * 256;
    const u32 lt = lanemask_lt();
    u32 peers_of[IPT];                 /* all ballots first: off the serial histogram chain */
#pragma unroll
    for (int j = 0; j < IPT; j++) peers_of[j] = match_digit((u32) (key[j] >> shift) & 255u);
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        const u32 digit = (u32) (key[j] >> shift) & 255u;
        const u32 peers = peers_of[j];
        const u32 leader = __ffs(peers) - 1;
        u32 c = 0;
        if (lane == leader) {
            c = my_hist[digit];
            my_hist[digit] = c + __popc(peers);
        }
        c = __shfl_sync(FULL_MASK, c, leader);
        rank[j] = c + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();

    /* ---- per digit: exclusive scan over warps, publish the tile count */
    u32 cnt_full = 0, cnt_valid = 0;
    if (tid < 256) {
        u32 c[WARPS];
#pragma unroll
        for (int w = 0; w < WARPS; w++) c[w] = s_whist[w * 256 + tid];
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
            s_whist[w * 256 + tid] = run;
            run += c[w];
        }
        cnt_full = run;
        cnt_valid = run;
        if (tid == 255) cnt_valid -= ((u32) TILE - valid);
        lookback_publish_partial(lb, tile, tid, cnt_valid);
        /* digit scan, warp part */
        u32 incl = cnt_full;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 y = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_misc[1 + warp] = incl;
        cnt_full = incl - cnt_full;  /* exclusive within the warp of digits */
    }
    __syncthreads();
    if (tid < 256) {
        u32 add = 0;
        for (u32 w = 0; w < warp; w++) add += s_misc[1 + w];
        const u32 local = cnt_full + add;      /* first slot of digit tid in the sorted tile */
        s_local[tid] = local;
#pragma unroll
        for (int w = 0; w < WARPS; w++) s_whist[w * 256 + tid] += local;
    }
    __syncthreads();

    /* ---- scatter keys into tile-sorted order in shared memory */
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        const u32 digit = (u32) (key[j] >> shift) & 255u;
        rank[j] += my_hist[digit];
        s_keys[rank[j]] = key[j];
    }

    /* ---- values: issue the loads now so they overlap the look-back spin */
    u32 val[IPT];
    if (IOTA) {
#pragma unroll
        for (int j = 0; j < IPT; j++) val[j] = wbase + j * 32;
    } else if (valid == (u32) TILE) {
#pragma unroll
        for (int j = 0; j < IPT; j++) val[j] = vin[wbase + j * 32];
    } else {
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            const u32 pos = wbase + j * 32;
            val[j] = pos < n ? vin[pos] : 0u;
        }
    }
    __syncthreads();  /* all ranks read from s_whist: it may now be reused as s_vals */

    /* ---- decoupled look-back: exclusive count of my digit over earlier tiles.
     * Four predecessors are polled per round trip (the walk is latency bound: one L2
     * access per predecessor otherwise). */
    if (tid < 256) {
        u32 excl = 0;
        if (tile > 0) {
            excl = lookback_exclusive(lb, tile, tid);
            st_relaxed_u32(&lb.tiles[(size_t) tile * 256 + tid], LB_INCL | (excl + cnt_valid));
        }
        s_gofs[tid] = bins[tid] + excl - s_local[tid];
    }
#pragma unroll
    for (int j = 0; j < IPT; j++) s_vals[rank[j]] = val[j];
    __syncthreads();

    /* ---- coalesced stores: consecutive threads write consecutive addresses of a run */
#pragma unroll
    for (int k = 0; k < IPT; k++) {
        const u32 s = tid + k * THREADS;
        if (s < valid) {
            const u64 kk = s_keys[s];
            const u32 digit = (u32) (kk >> shift) & 255u;
            const u32 g = s_gofs[digit] + s;
            kout[g] = kk;
            vout[g] = s_vals[s];
        }
    }
}

typedef SweepCfg<MPSK_SWEEP_THREADS, MPSK_SWEEP_IPT> TheSweep;

extern "C" size_t mpsk_onesweep_tile_items(void) { return TheSweep::TILE; }

extern "C" size_t mpsk_onesweep_scratch_bytes(size_t n)
{
    /* enough for the smallest tile of any pass flavour (record passes use 3072) */
    const size_t tile = 2048;
    const size_t ntiles = (n + tile - 1) / tile;
    return lookback_words(ntiles) * sizeof(u32);
}

extern "C" int mpsk_onesweep_pass(const uint64_t * kin, const uint32_t * vin,
        uint64_t * kout, uint32_t * vout, size_t n, int shift,
        const uint32_t * bins, void * scratch, mpsk_stream_t stream_)
{
    if (n == 0) return 0;
    if (n > MPSK_MAX_ITEMS) return (int) cudaErrorInvalidValue;
    cudaStream_t stream = (cudaStream_t) stream_;
    const size_t ntiles = (n + TheSweep::TILE - 1) / TheSweep::TILE;
    cudaError_t e = cudaMemsetAsync(scratch, 0, lookback_words(ntiles) * sizeof(u32), stream);
    if (e != cudaSuccess) return (int) e;
    u32 * ticket = (u32 *) scratch;
    LookbackBufs lb;
    lb.tiles = ticket + 64;
    lb.blktotal = lb.tiles + ntiles * 256;
    lb.blkincl = lb.blktotal + ((ntiles + LB_BLOCK - 1) / LB_BLOCK) * 256;
    /* the attribute is per device: set it on every launch (local groups span devices) */
    if (vin == NULL) {
        auto kern = onesweep_kernel<MPSK_SWEEP_THREADS, MPSK_SWEEP_IPT, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TheSweep::SMEM);
        if (e != cudaSuccess) return (int) e;
        kern<<<(unsigned) ntiles, MPSK_SWEEP_THREADS, TheSweep::SMEM, stream>>>(
            (const u64 *) kin, vin, (u64 *) kout, vout, (u32) n, (u32) shift, bins, lb, ticket);
    } else {
        auto kern = onesweep_kernel<MPSK_SWEEP_THREADS, MPSK_SWEEP_IPT, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TheSweep::SMEM);
        if (e != cudaSuccess) return (int) e;
        kern<<<(unsigned) ntiles, MPSK_SWEEP_THREADS, TheSweep::SMEM, stream>>>(
            (const u64 *) kin, vin, (u64 *) kout, vout, (u32) n, (u32) shift, bins, lb, ticket);
    }
    CUDA_LAUNCH_CHECK();
    return 0;
}

/* ------------------------------------------------------------------------- */
/* onesweep pass over whole 16-byte records {u64 key, u64 payload} (either order) */
/*
 * When a record is nothing but its 8-byte key and 8 more bytes, carrying the record
 * through the passes (32 B of HBM traffic per record and pass) is cheaper than
 * sorting (key, index) pairs (24 B) and gathering afterwards: a random 16-byte read
 * costs a whole 128-byte DRAM line on B200 (profiles/r01_gather_probe.log), i.e. the
 * gather alone moves 148 B per record. Same algorithm as onesweep_kernel; the key is
 * read in place (low or high half, sign flip applied on the fly), items move as uint4.
 */
#ifndef MPSK_REC_THREADS
#define MPSK_REC_THREADS 384
#endif
#ifndef MPSK_REC_IPT
#define MPSK_REC_IPT 8
#endif
#ifndef MPSK_REC_MINBLOCKS
#define MPSK_REC_MINBLOCKS 3
#endif
#ifndef MPSK_REC8_IPT
#define MPSK_REC8_IPT 12
#endif
#ifndef MPSK_REC_TMA_STORE
#define MPSK_REC_TMA_STORE 1
#endif

template <int THREADS, int IPT, int ITEMBYTES>
struct RecCfg {
    static constexpr int TILE = THREADS * IPT;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int SMEM = TILE * ITEMBYTES + WARPS * 256 * 4 + 256 * 4 * 2 + 64;
};

/* an item is a whole record: uint4 = {u64, u64} with the key in either half, or a bare u64 key */
__device__ __forceinline__ u32 rec_digit(const uint4 & it, u32 khi, u64 flip, u32 shift)
{
    const u64 k = (khi ? (((u64) it.w << 32) | it.z) : (((u64) it.y << 32) | it.x)) ^ flip;
    return (u32) (k >> shift) & 255u;
}
__device__ __forceinline__ u32 rec_digit(const u64 & it, u32 khi, u64 flip, u32 shift)
{
    (void) khi;
    return (u32) ((it ^ flip) >> shift) & 255u;
}
__device__ __forceinline__ void rec_pad(uint4 & it, u64 padk) { it = make_uint4((u32) padk, (u32) (padk >> 32), (u32) padk, (u32) (padk >> 32)); }
__device__ __forceinline__ void rec_pad(u64 & it, u64 padk) { it = padk; }

template <int THREADS, int IPT, typename ITEM>
__global__ void __launch_bounds__(THREADS, MPSK_REC_MINBLOCKS)
onesweep_rec_kernel(const ITEM * __restrict__ in, ITEM * __restrict__ out,
                      u32 n, u32 shift, u32 khi, u64 flip, const u32 * __restrict__ bins,
                      LookbackBufs lb, u32 * ticket)
{
    typedef RecCfg<THREADS, IPT, (int) sizeof(ITEM)> Cfg;
    constexpr int TILE = Cfg::TILE;
    constexpr int WARPS = Cfg::WARPS;
    static_assert(THREADS >= 256, "one thread per digit needs >= 256 threads");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    ITEM * s_items = (ITEM *) smem_raw;
    u32 * s_whist = (u32 *) (smem_raw + TILE * sizeof(ITEM));     /* [WARPS][256] */
    u32 * s_local = s_whist + WARPS * 256;
    u32 * s_gofs = s_local + 256;
    u32 * s_misc = s_gofs + 256;

    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31u;
    const u32 warp = tid >> 5;

    if (tid == 0) s_misc[0] = atomicAdd(ticket, 1u);
    for (u32 i = tid; i < WARPS * 256; i += THREADS) s_whist[i] = 0;
    __syncthreads();

    const u32 tile = s_misc[0];
    const u32 tile_base = tile * (u32) TILE;
    const u32 remaining = n - tile_base;
    const u32 valid = remaining < (u32) TILE ? remaining : (u32) TILE;
    const u32 wbase = tile_base + warp * (IPT * 32) + lane;

    /* ---- load records, warp-striped: each load instruction covers 512 contiguous bytes */
    ITEM it[IPT];
    if (valid == (u32) TILE) {
#pragma unroll
        for (int j = 0; j < IPT; j++) it[j] = in[wbase + j * 32];
    } else {
        /* padding must rank last in bin 255: (key ^ flip) == ~0 */
        ITEM pad;
        rec_pad(pad, ~flip);
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            const u32 pos = wbase + j * 32;
            it[j] = pos < n ? in[pos] : pad;
        }
    }

    /* ---- rank inside (warp, digit). The eight ballots of every row are independent of
     * the serial histogram update below: issue them all first so their latency overlaps. */
    u32 rank[IPT];
    u32 * my_hist = s_whist + warp * 256;
    const u32 lt = lanemask_lt();
    u32 peers_of[IPT];
#pragma unroll
    for (int j = 0; j < IPT; j++) peers_of[j] = match_digit(rec_digit(it[j], khi, flip, shift));
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        const u32 digit = rec_digit(it[j], khi, flip, shift);
        const u32 peers = peers_of[j];
        const u32 leader = __ffs(peers) - 1;
        u32 c = 0;
        if (lane == leader) {
            c = my_hist[digit];
            my_hist[digit] = c + __popc(peers);
        }
        c = __shfl_sync(FULL_MASK, c, leader);
        rank[j] = c + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();

    /* ---- per digit: exclusive scan over warps, publish the tile count */
    u32 cnt_full = 0, cnt_valid = 0;
    if (tid < 256) {
        u32 c[WARPS];
#pragma unroll
        for (int w = 0; w < WARPS; w++) c[w] = s_whist[w * 256 + tid];
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
            s_whist[w * 256 + tid] = run;
            run += c[w];
        }
        cnt_full = run;
        cnt_valid = run;
        if (tid == 255) cnt_valid -= ((u32) TILE - valid);
        lookback_publish_partial(lb, tile, tid, cnt_valid);
        u32 incl = cnt_full;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 y = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_misc[1 + warp] = incl;
        cnt_full = incl - cnt_full;
    }
    __syncthreads();
    if (tid < 256) {
        u32 add = 0;
        for (u32 w = 0; w < warp; w++) add += s_misc[1 + w];
        const u32 local = cnt_full + add;
        s_local[tid] = local;
#pragma unroll
        for (int w = 0; w < WARPS; w++) s_whist[w * 256 + tid] += local;
    }
    __syncthreads();

    /* ---- scatter records into tile-sorted order in shared memory */
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        const u32 digit = rec_digit(it[j], khi, flip, shift);
        s_items[rank[j] + my_hist[digit]] = it[j];
    }
#if MPSK_REC_TMA_STORE
    /* the staged tile is read by the bulk-copy engine below: make the generic-proxy
     * shared-memory writes visible to the async proxy */
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif

    /* ---- decoupled look-back (see onesweep_kernel) */
    u32 excl = 0;
    if (tid < 256) {
        if (tile > 0) {
            excl = lookback_exclusive(lb, tile, tid);
            st_relaxed_u32(&lb.tiles[(size_t) tile * 256 + tid], LB_INCL | (excl + cnt_valid));
        }
        s_gofs[tid] = bins[tid] + excl - s_local[tid];
    }
    __syncthreads();

#if MPSK_REC_TMA_STORE
    /* ---- one bulk copy (TMA, cp.async.bulk shared -> global) per digit run: the run of
     * digit d is contiguous both in the staged tile and in the output, a multiple of the
     * record size long and 16-byte aligned on both sides. 256 threads issue 256 copies;
     * nobody executes a per-record store loop. */
    if (sizeof(ITEM) == 16) {
        if (tid < 256) {
            if (cnt_valid) {
                const u32 local = s_local[tid];
                ITEM * dst = out + (bins[tid] + excl);
                const u32 src = (u32) __cvta_generic_to_shared(&s_items[local]);
                const u32 bytes = cnt_valid * (u32) sizeof(ITEM);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             :: "l"(dst), "r"(src), "r"(bytes) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            /* shared memory must stay valid until the engine has read it */
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else
#endif
    {
        /* ---- coalesced stores of the digit runs (8-byte items: runs are not multiples of 16 bytes) */
#pragma unroll
        for (int k = 0; k < IPT; k++) {
            const u32 s = tid + k * THREADS;
            if (s < valid) {
                const ITEM v = s_items[s];
                out[s_gofs[rec_digit(v, khi, flip, shift)] + s] = v;
            }
        }
    }
}

/*
 * CANDIDATE, not the default and not yet measured in this form (MPSK_REC_PERSIST=1 selects it):
 * persistent form of the record pass. The grid is one wave of resident CTAs and every CTA
 * takes tiles by ticket until none are left, so that
 *   - the next tile's records are loaded right after the current tile has been scattered to
 *     shared memory (the item registers are free then): their HBM latency overlaps the
 *     look-back and the bulk stores (ncu of the one-tile-per-CTA kernel: 12 % of the warp
 *     samples wait for the tile's own loads, profiles/r01_ncu_rec16_s3_summary.txt);
 *   - the wait for the bulk-copy engine to have read the staged tile moves from right after
 *     the copies to just before the NEXT tile is scattered, two barriers later.
 * A first version that waited for the bulk reads immediately was bit-exact but 9 % slower than
 * the default kernel (profiles/r01_sweep5_rec_shapes.log). Tickets are handed out in tile order
 * to running CTAs, so every tile a look-back waits for is held by a CTA that never waits for a
 * later tile.
 */
#ifndef MPSK_REC_PERSIST
#define MPSK_REC_PERSIST 0
#endif
#ifndef MPSK_REC_PERSIST_PREFETCH
#define MPSK_REC_PERSIST_PREFETCH 1
#endif
#if MPSK_REC_PERSIST
template <int IPT, typename ITEM>
__device__ __forceinline__ void rec_load_tile(ITEM (&it)[IPT], const ITEM * __restrict__ in, u32 n, u32 tile_base,
                                              u32 tile_items, u32 woff, u64 flip)
{
    const u32 wbase = tile_base + woff;
    if (n - tile_base >= tile_items) {
#pragma unroll
        for (int j = 0; j < IPT; j++) it[j] = in[wbase + j * 32];
    } else {
        ITEM pad;
        rec_pad(pad, ~flip);               /* padding ranks last in bin 255: (key ^ flip) == ~0 */
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            const u32 pos = wbase + j * 32;
            it[j] = pos < n ? in[pos] : pad;
        }
    }
}

template <int THREADS, int IPT, typename ITEM>
__global__ void __launch_bounds__(THREADS, MPSK_REC_MINBLOCKS)
onesweep_rec_persist_kernel(const ITEM * __restrict__ in, ITEM * __restrict__ out,
                            u32 n, u32 ntiles, u32 shift, u32 khi, u64 flip, const u32 * __restrict__ bins,
                            LookbackBufs lb, u32 * ticket)
{
    typedef RecCfg<THREADS, IPT, (int) sizeof(ITEM)> Cfg;
    constexpr int TILE = Cfg::TILE;
    constexpr int WARPS = Cfg::WARPS;
    constexpr bool BULK = sizeof(ITEM) == 16;
    static_assert(THREADS >= 288, "needs the 256 digit threads plus one more warp");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    ITEM * s_items = (ITEM *) smem_raw;
    u32 * s_whist = (u32 *) (smem_raw + TILE * sizeof(ITEM));     /* [WARPS][256] */
    u32 * s_local = s_whist + WARPS * 256;
    u32 * s_gofs = s_local + 256;
    u32 * s_misc = s_gofs + 256;       /* [0] first tile, [1..8] digit-scan warp totals, [12] next tile */

    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31u;
    const u32 warp = tid >> 5;
    const u32 woff = warp * (IPT * 32) + lane;
    u32 * my_hist = s_whist + warp * 256;
    const u32 lt = lanemask_lt();

    if (tid == 0) s_misc[0] = atomicAdd(ticket, 1u);
#pragma unroll
    for (int k = 0; k < 8; k++) my_hist[lane + 32 * k] = 0;
    __syncthreads();
    u32 tile = s_misc[0];
    if (tile >= ntiles) return;

    ITEM it[IPT];
    rec_load_tile<IPT, ITEM>(it, in, n, tile * (u32) TILE, (u32) TILE, woff, flip);

    for (;;) {
        const u32 tile_base = tile * (u32) TILE;
        const u32 remaining = n - tile_base;
        const u32 valid = remaining < (u32) TILE ? remaining : (u32) TILE;

        /* ---- rank inside (warp, digit): all ballots first, then the serial histogram chain */
        u32 rank[IPT];
        {
            u32 peers_of[IPT];
#pragma unroll
            for (int j = 0; j < IPT; j++) peers_of[j] = match_digit(rec_digit(it[j], khi, flip, shift));
#pragma unroll
            for (int j = 0; j < IPT; j++) {
                const u32 digit = rec_digit(it[j], khi, flip, shift);
                const u32 peers = peers_of[j];
                const u32 leader = __ffs(peers) - 1;
                u32 c = 0;
                if (lane == leader) {
                    c = my_hist[digit];
                    my_hist[digit] = c + __popc(peers);
                }
                c = __shfl_sync(FULL_MASK, c, leader);
                rank[j] = c + __popc(peers & lt);
                __syncwarp();
            }
        }
        __syncthreads();                                                   /* (A) */

        /* ---- per digit: exclusive scan over warps, publish the tile count */
        u32 cnt_full = 0, cnt_valid = 0;
        if (tid < 256) {
            u32 c[WARPS];
#pragma unroll
            for (int w = 0; w < WARPS; w++) c[w] = s_whist[w * 256 + tid];
            u32 run = 0;
#pragma unroll
            for (int w = 0; w < WARPS; w++) {
                s_whist[w * 256 + tid] = run;
                run += c[w];
            }
            cnt_full = run;
            cnt_valid = run;
            if (tid == 255) cnt_valid -= ((u32) TILE - valid);
            lookback_publish_partial(lb, tile, tid, cnt_valid);
            u32 incl = cnt_full;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 y = __shfl_up_sync(FULL_MASK, incl, o);
                if (lane >= o) incl += y;
            }
            if (lane == 31) s_misc[1 + warp] = incl;
            cnt_full = incl - cnt_full;
        } else if (tid == THREADS - 1) {
            /* a thread of a warp that idles through the digit phases takes the next ticket */
            s_misc[12] = atomicAdd(ticket, 1u);
        }
        __syncthreads();                                                   /* (B) */
        if (tid < 256) {
            u32 add = 0;
            for (u32 w = 0; w < warp; w++) add += s_misc[1 + w];
            const u32 local = cnt_full + add;
            s_local[tid] = local;
#pragma unroll
            for (int w = 0; w < WARPS; w++) s_whist[w * 256 + tid] += local;
            /* the bulk copies of the PREVIOUS tile (issued by this thread) must have read the
             * staged tile before anybody scatters into it again, i.e. before barrier (C) */
            if (BULK) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();                                                   /* (C) */

        /* ---- scatter records into tile-sorted order in shared memory */
#pragma unroll
        for (int j = 0; j < IPT; j++) {
            const u32 digit = rec_digit(it[j], khi, flip, shift);
            s_items[rank[j] + my_hist[digit]] = it[j];
        }
        if (BULK) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

        /* ---- the item registers are free: start loading the next tile
         * (MPSK_REC_PERSIST_PREFETCH=0 loads it after the stores instead: isolates the effect) */
        const u32 next = s_misc[12];
#if MPSK_REC_PERSIST_PREFETCH
        if (next < ntiles) rec_load_tile<IPT, ITEM>(it, in, n, next * (u32) TILE, (u32) TILE, woff, flip);
#endif

        /* ---- decoupled look-back */
        u32 excl = 0;
        if (tid < 256) {
            if (tile > 0) {
                excl = lookback_exclusive(lb, tile, tid);
                st_relaxed_u32(&lb.tiles[(size_t) tile * 256 + tid], LB_INCL | (excl + cnt_valid));
            }
            if (!BULK) s_gofs[tid] = bins[tid] + excl - s_local[tid];
        }
        __syncthreads();                                                   /* (D) */

        if (BULK) {
            /* one bulk copy (cp.async.bulk shared -> global) per digit run; not waited for here */
            if (tid < 256) {
                if (cnt_valid) {
                    const u32 local = s_local[tid];
                    ITEM * dst = out + (bins[tid] + excl);
                    const u32 src = (u32) __cvta_generic_to_shared(&s_items[local]);
                    const u32 bytes = cnt_valid * (u32) sizeof(ITEM);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 :: "l"(dst), "r"(src), "r"(bytes) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
#pragma unroll
            for (int k = 0; k < IPT; k++) {
                const u32 s = tid + k * THREADS;
                if (s < valid) {
                    const ITEM v = s_items[s];
                    out[s_gofs[rec_digit(v, khi, flip, shift)] + s] = v;
                }
            }
        }
        if (next >= ntiles) break;
        tile = next;
#if !MPSK_REC_PERSIST_PREFETCH
        rec_load_tile<IPT, ITEM>(it, in, n, tile * (u32) TILE, (u32) TILE, woff, flip);
#endif
        /* every warp resets its own histogram: past (D) nobody reads it any more, and the
         * other warps touch it again only after (A) of the next tile. (8-byte items: the store
         * loop above reads s_items and s_gofs; both are rewritten only after (B)/(C).) */
#pragma unroll
        for (int k = 0; k < 8; k++) my_hist[lane + 32 * k] = 0;
        __syncwarp();
    }
    /* shared memory must stay valid until the engine has read the last tile */
    if (BULK && tid < 256) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
#endif

template <typename ITEM>
static int launch_rec_pass(const void * in, void * out, size_t n, int shift, int key_in_high, uint64_t flip,
                           const uint32_t * bins, void * scratch, cudaStream_t stream)
{
#if MPSK_REC_PERSIST
    {
        constexpr int IPT = sizeof(ITEM) == 8 ? MPSK_REC8_IPT : MPSK_REC_IPT;
        typedef RecCfg<MPSK_REC_THREADS, IPT, (int) sizeof(ITEM)> Cfg;
        const size_t ntiles = (n + Cfg::TILE - 1) / Cfg::TILE;
        cudaError_t e = cudaMemsetAsync(scratch, 0, lookback_words(ntiles) * sizeof(u32), stream);
        if (e != cudaSuccess) return (int) e;
        u32 * ticket = (u32 *) scratch;
        LookbackBufs lb;
        lb.tiles = ticket + 64;
        lb.blktotal = lb.tiles + ntiles * 256;
        lb.blkincl = lb.blktotal + ((ntiles + LB_BLOCK - 1) / LB_BLOCK) * 256;
        auto kern = onesweep_rec_persist_kernel<MPSK_REC_THREADS, IPT, ITEM>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return (int) e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, MPSK_REC_THREADS, Cfg::SMEM);
        if (e != cudaSuccess) return (int) e;
        if (per_sm < 1) per_sm = 1;
        size_t grid = (size_t) per_sm * (size_t) num_sms();
        if (grid > ntiles) grid = ntiles;
        kern<<<(unsigned) grid, MPSK_REC_THREADS, Cfg::SMEM, stream>>>(
            (const ITEM *) in, (ITEM *) out, (u32) n, (u32) ntiles, (u32) shift, key_in_high ? 1u : 0u, (u64) flip,
            bins, lb, ticket);
        CUDA_LAUNCH_CHECK();
        return 0;
    }
#else
    /* bare 8-byte keys: 12 per thread is the best of the shapes tried (profiles/r01_sweep5_rec_shapes.log) */
    constexpr int IPT = sizeof(ITEM) == 8 ? MPSK_REC8_IPT : MPSK_REC_IPT;
    typedef RecCfg<MPSK_REC_THREADS, IPT, (int) sizeof(ITEM)> Cfg;
    const size_t ntiles = (n + Cfg::TILE - 1) / Cfg::TILE;
    cudaError_t e = cudaMemsetAsync(scratch, 0, lookback_words(ntiles) * sizeof(u32), stream);
    if (e != cudaSuccess) return (int) e;
    u32 * ticket = (u32 *) scratch;
    LookbackBufs lb;
    lb.tiles = ticket + 64;
    lb.blktotal = lb.tiles + ntiles * 256;
    lb.blkincl = lb.blktotal + ((ntiles + LB_BLOCK - 1) / LB_BLOCK) * 256;
    auto kern = onesweep_rec_kernel<MPSK_REC_THREADS, IPT, ITEM>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return (int) e;
    kern<<<(unsigned) ntiles, MPSK_REC_THREADS, Cfg::SMEM, stream>>>(
        (const ITEM *) in, (ITEM *) out, (u32) n, (u32) shift, key_in_high ? 1u : 0u, (u64) flip,
        bins, lb, ticket);
    CUDA_LAUNCH_CHECK();
    return 0;
#endif
}

extern "C" int mpsk_onesweep_pass_rec(const void * in, void * out, size_t n, size_t elsize, int shift,
        int key_in_high, uint64_t flip, const uint32_t * bins, void * scratch, mpsk_stream_t stream_)
{
    if (n == 0) return 0;
    if (n > MPSK_MAX_ITEMS) return (int) cudaErrorInvalidValue;
    cudaStream_t stream = (cudaStream_t) stream_;
    if (elsize == 16) return launch_rec_pass<uint4>(in, out, n, shift, key_in_high, flip, bins, scratch, stream);
    if (elsize == 8) return launch_rec_pass<u64>(in, out, n, shift, 0, flip, bins, scratch, stream);
    return (int) cudaErrorInvalidValue;
}

/* ------------------------------------------------------------------------- */
/* hybrid sort: LSD over the four most significant non-constant digits, then the   */
/* runs of equal high part are ordered by their low part in place                  */
/*
 * For keys whose high 32 significant bits are (nearly) distinct -- random 64-bit ids,
 * hashes -- four passes already put almost every record in its final place: what is
 * left are short runs of records that agree in the high part. Inside a run records
 * are still in input order (the passes are stable), so ordering a run stably by the
 * low part gives exactly the order of the full eight-pass sort.
 *
 * fixup_rec_kernel: one CTA per tile of FIX_T records (+ FIX_HALO look-ahead).
 * A run belongs to the tile that holds its head. Runs of 2..FIX_HALO records are
 * ranked by counting (O(L^2), L is tiny) and rewritten in place; longer runs are
 * appended to a work list and sorted by the host with ordinary passes.
 * Safe in place: a CTA rewrites only runs it owns, and what other CTAs read of those
 * records (the high part, for head detection) does not change when a run is permuted.
 */
#define FIX_T 2048
#define FIX_HALO 256
#define FIX_THREADS 256

__device__ __forceinline__ u64 rec_key(const uint4 & it, u32 khi, u64 flip)
{
    return (khi ? (((u64) it.w << 32) | it.z) : (((u64) it.y << 32) | it.x)) ^ flip;
}

template <typename ITEM, bool KHI>
__global__ void __launch_bounds__(FIX_THREADS)
fixup_rec_kernel(ITEM * __restrict__ recs, u32 n, u64 flip, u32 lobits,
                   u32 * __restrict__ worklist, u32 * __restrict__ nwork, u32 cap)
{
    constexpr int CAP = FIX_T + FIX_HALO;
    constexpr int WORDS = (CAP + 31) / 32 + 1;
    constexpr int NLD = CAP / FIX_THREADS;
    static_assert(CAP % FIX_THREADS == 0, "tile + halo must be a multiple of the block size");
    /* only the keys are staged: the few records that move are re-read from global */
    __shared__ u64 s_key[CAP + 1];            /* [0] = key of the record before the tile */
    __shared__ u32 s_head[WORDS];

    const u32 tid = threadIdx.x;
    const size_t t0 = (size_t) blockIdx.x * FIX_T;
    const u32 avail = (u32) ((size_t) n - t0);
    const u32 cnt = avail < (u32) CAP ? avail : (u32) CAP;
    const bool at_end = (t0 + cnt == n);
    const u64 lomask = lobits >= 64 ? ~0ULL : ((1ULL << lobits) - 1ULL);
    constexpr u32 W = sizeof(ITEM) / 8;                            /* u64 words per record */
    const u64 * keys = (const u64 *) recs + (KHI ? 1 : 0);       /* key of record i at keys[W*i] */

    {
        u64 tmp[NLD];
#pragma unroll
        for (int k = 0; k < NLD; k++) {
            const u32 i = tid + k * FIX_THREADS;
            if (i < cnt) tmp[k] = keys[W * (t0 + i)];
        }
#pragma unroll
        for (int k = 0; k < NLD; k++) {
            const u32 i = tid + k * FIX_THREADS;
            if (i < cnt) s_key[i + 1] = tmp[k] ^ flip;
        }
        if (tid == 0) s_key[0] = t0 ? (keys[W * (t0 - 1)] ^ flip) : 0ULL;
    }
    __syncthreads();
    /* head flags: the high part differs from the predecessor's. A warp handles 32
     * consecutive positions per round, so one ballot is one word of the bit map. */
#pragma unroll
    for (int k = 0; k < (WORDS * 32 + FIX_THREADS - 1) / FIX_THREADS; k++) {
        const u32 i = tid + k * FIX_THREADS;
        if (i < (u32) WORDS * 32) {
            bool head = false;
            if (i == cnt) head = at_end;                      /* sentinel: the data ends here */
            else if (i < cnt) head = ((s_key[i] >> lobits) != (s_key[i + 1] >> lobits)) || (i == 0 && t0 == 0);
            const u32 word = __ballot_sync(FULL_MASK, head);
            if ((tid & 31) == 0) s_head[i >> 5] = word;
        }
    }
    __syncthreads();
    /* ---- compact the positions that are NOT a run of their own (6 % for random keys):
     * the expensive part below then runs with full warps */
    __shared__ u32 s_list[CAP];
    __shared__ u32 s_nlist;
    if (tid == 0) s_nlist = 0;
    __syncthreads();
    /* a warp's 32 positions of round k are exactly the bits of head word (warp + 8k):
     * position i is a run of its own when bits i and i+1 are both set, so the whole
     * row is decided by two broadcast loads and a few word operations */
#pragma unroll
    for (int k = 0; k < NLD; k++) {
        const u32 w = (tid >> 5) + k * (FIX_THREADS / 32);
        const u32 H = s_head[w], Hn = s_head[w + 1];
        const u32 single = H & ((H >> 1) | (Hn << 31));
        const u32 first = w * 32;
        const u32 inside = first >= cnt ? 0u : (cnt - first >= 32 ? 0xffffffffu : ((1u << (cnt - first)) - 1u));
        const u32 votes = ~single & inside;
        if (votes) {
            u32 base = 0;
            if ((tid & 31) == 0) base = atomicAdd(&s_nlist, (u32) __popc(votes));
            base = __shfl_sync(FULL_MASK, base, 0);
            if ((votes >> (tid & 31)) & 1u) s_list[base + __popc(votes & lanemask_lt())] = first + (tid & 31);
        }
    }
    __syncthreads();
    const u32 nlist = s_nlist;
    ITEM moved[NLD];
    u32 tgts[NLD];
#pragma unroll
    for (int k = 0; k < NLD; k++) {
        tgts[k] = 0xffffffffu;
        const u32 e = tid + k * FIX_THREADS;
        if (e >= nlist) continue;
        const u32 i = s_list[e];
        /* run start: last head at or before i */
        int w = (int) (i >> 5);
        u32 bits = s_head[w] & (0xffffffffu >> (31 - (i & 31)));
        while (bits == 0 && w > 0) { w--; bits = s_head[w]; }
        if (bits == 0) continue;                          /* continuation of a run owned by an earlier tile */
        const u32 rs = (u32) w * 32 + (31 - __clz(bits));
        if (rs >= (u32) FIX_T) continue;                  /* head lies in the look-ahead: the next tile owns it */
        /* run end: first head after i (the sentinel counts) */
        u32 w2 = (i + 1) >> 5;
        u32 b2 = s_head[w2] & (0xffffffffu << ((i + 1) & 31));
        while (b2 == 0 && w2 + 1 < (u32) WORDS && (w2 + 1) * 32 <= cnt + 31) { w2++; b2 = s_head[w2]; }
        const u32 re = b2 ? (w2 * 32 + (__ffs(b2) - 1)) : 0xffffffffu;
        if (re == 0xffffffffu || re > cnt || re - rs > (u32) FIX_HALO) {
            /* too long for this kernel: the run head reports it */
            if (i == rs) {
                const u32 slot = atomicAdd(nwork, 1u);
                if (slot < cap) worklist[slot] = (u32) (t0 + rs);
            }
            continue;
        }
        if (re - rs < 2) continue;
        const u64 mine = s_key[i + 1] & lomask;
        u32 rank = 0;
        for (u32 j = rs; j < re; j++) {
            const u64 other = s_key[j + 1] & lomask;
            rank += (other < mine) || (other == mine && j < i);
        }
        if (rs + rank != i) {
            tgts[k] = rs + rank;
            moved[k] = recs[t0 + i];                      /* read before anyone of this CTA writes */
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NLD; k++)
        if (tgts[k] != 0xffffffffu) recs[t0 + tgts[k]] = moved[k];
}

/* extent of every long run on the work list: first index whose high part differs */
__device__ __forceinline__ u64 rec_key_at(const u64 * words, size_t i, u32 W, u32 khi, u64 flip)
{
    return words[W * i + khi] ^ flip;
}

__global__ void fixup_extent_kernel(const u64 * __restrict__ recs, u32 W, u32 n, u32 khi, u64 flip, u32 lobits,
                                    const u32 * __restrict__ worklist, u32 nwork, u32 * __restrict__ lengths)
{
    const u32 e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nwork) return;
    const u32 start = worklist[e];
    const u64 hi = rec_key_at(recs, start, W, khi, flip) >> lobits;
    u32 lo = start + 1, hiidx = n;                       /* keys are sorted by the high part */
    while (lo < hiidx) {
        const u32 mid = lo + ((hiidx - lo) >> 1);
        if ((rec_key_at(recs, mid, W, khi, flip) >> lobits) <= hi) lo = mid + 1; else hiidx = mid;
    }
    lengths[e] = lo - start;
}

extern "C" int mpsk_fixup_rec(void * recs, size_t n, size_t elsize, int key_in_high, uint64_t flip, uint32_t lobits,
        uint32_t * worklist, uint32_t * nwork, uint32_t cap, mpsk_stream_t stream)
{
    if (n == 0) return 0;
    const size_t tiles = (n + FIX_T - 1) / FIX_T;
    if (elsize == 8)
        fixup_rec_kernel<u64, false><<<(unsigned) tiles, FIX_THREADS, 0, (cudaStream_t) stream>>>(
            (u64 *) recs, (u32) n, (u64) flip, lobits, worklist, nwork, cap);
    else if (key_in_high)
        fixup_rec_kernel<uint4, true><<<(unsigned) tiles, FIX_THREADS, 0, (cudaStream_t) stream>>>(
            (uint4 *) recs, (u32) n, (u64) flip, lobits, worklist, nwork, cap);
    else
        fixup_rec_kernel<uint4, false><<<(unsigned) tiles, FIX_THREADS, 0, (cudaStream_t) stream>>>(
            (uint4 *) recs, (u32) n, (u64) flip, lobits, worklist, nwork, cap);
    CUDA_LAUNCH_CHECK();
    return 0;
}

extern "C" int mpsk_fixup_extents(const void * recs, size_t n, size_t elsize, int key_in_high, uint64_t flip, uint32_t lobits,
        const uint32_t * worklist, uint32_t nwork, uint32_t * lengths, mpsk_stream_t stream)
{
    if (nwork == 0) return 0;
    fixup_extent_kernel<<<(nwork + 63) / 64, 64, 0, (cudaStream_t) stream>>>(
        (const u64 *) recs, (u32) (elsize / 8), (u32) n, (key_in_high && elsize == 16) ? 1u : 0u, (u64) flip, lobits, worklist, nwork, lengths);
    CUDA_LAUNCH_CHECK();
    return 0;
}

/* predictor: the high parts of `s` evenly spaced records, as bare u64 "records" */
__global__ void sample_prefix_kernel(const u64 * __restrict__ recs, u32 W, size_t n, u32 s, u32 khi, u64 flip, u32 lobits,
                                     u64 * __restrict__ out)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s) return;
    const size_t pos = (size_t) (((unsigned __int128) i * n) / s);
    out[i] = rec_key_at(recs, pos, W, khi, flip) >> lobits;
}

/* number of equal PAIRS in a sorted array: sum over values of k(k-1)/2 */
__global__ void count_equal_pairs_kernel(const u64 * __restrict__ sorted, u32 s, u64 * __restrict__ count)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    u64 pairs = 0;
    if (i < s) {
        const u64 v = sorted[i];
        u32 lo = 0, hi = i;                               /* first index holding v */
        while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (sorted[mid] < v) lo = mid + 1; else hi = mid; }
        pairs = i - lo;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pairs += __shfl_xor_sync(FULL_MASK, pairs, o);
    if ((threadIdx.x & 31) == 0 && pairs) atomicAdd(count, pairs);
}

extern "C" int mpsk_sample_prefix_rec(const void * recs, size_t n, size_t elsize, uint32_t s, int key_in_high, uint64_t flip,
        uint32_t lobits, uint64_t * out, mpsk_stream_t stream)
{
    if (s == 0) return 0;
    sample_prefix_kernel<<<(s + 255) / 256, 256, 0, (cudaStream_t) stream>>>(
        (const u64 *) recs, (u32) (elsize / 8), n, s, (key_in_high && elsize == 16) ? 1u : 0u, (u64) flip, lobits, (u64 *) out);
    CUDA_LAUNCH_CHECK();
    return 0;
}

extern "C" int mpsk_count_equal_pairs(const uint64_t * sorted, uint32_t s, uint64_t * count, mpsk_stream_t stream)
{
    if (s == 0) return 0;
    count_equal_pairs_kernel<<<(s + 255) / 256, 256, 0, (cudaStream_t) stream>>>((const u64 *) sorted, s, (u64 *) count);
    CUDA_LAUNCH_CHECK();
    return 0;
}

/* ========================================================================= */
/* gathers                                                                   */
/* ========================================================================= */

__global__ void __launch_bounds__(256)
gather_u64_kernel(const u64 * __restrict__ src, const u32 * __restrict__ idx,
                  u64 * __restrict__ dst, size_t n)
{
    const size_t base = ((size_t) blockIdx.x * blockDim.x) * 4 + threadIdx.x;
    u32 ix[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const size_t i = base + (size_t) k * blockDim.x;
        ix[k] = i < n ? idx[i] : 0u;
    }
    u64 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const size_t i = base + (size_t) k * blockDim.x;
        v[k] = i < n ? src[ix[k]] : 0ULL;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const size_t i = base + (size_t) k * blockDim.x;
        if (i < n) dst[i] = v[k];
    }
}

extern "C" int mpsk_gather_u64(const uint64_t * src, const uint32_t * idx, uint64_t * dst,
        size_t n, mpsk_stream_t stream)
{
    if (n == 0) return 0;
    const size_t per_block = 256 * 4;
    const size_t blocks = (n + per_block - 1) / per_block;
    gather_u64_kernel<<<(unsigned) blocks, 256, 0, (cudaStream_t) stream>>>(
        (const u64 *) src, idx, (u64 *) dst, n);
    CUDA_LAUNCH_CHECK();
    return 0;
}

/*
 * K3 payload gather. A record of `elsize` bytes is moved by elsize/VEC lanes, each
 * moving one VEC-byte piece, so the lanes of one record read/write consecutive
 * addresses. Writes are fully coalesced (out is written in order); reads are one
 * random record each. UNROLL independent records per thread keep enough loads in
 * flight to cover the random-access latency.
 */
template <typename V, int UNROLL>
__global__ void __launch_bounds__(256)
gather_records_kernel(const V * __restrict__ base, const u32 * __restrict__ idx,
                      V * __restrict__ out, size_t n, u32 lpr /* lanes per record */)
{
    const size_t total = n * (size_t) lpr;
    const size_t t0 = ((size_t) blockIdx.x * blockDim.x) * UNROLL + threadIdx.x;
    size_t src[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; k++) {
        const size_t t = t0 + (size_t) k * blockDim.x;
        if (t < total) {
            const size_t rec = t / lpr;
            const u32 part = (u32) (t - rec * lpr);
            src[k] = (size_t) idx[rec] * lpr + part;
        } else {
            src[k] = 0;
        }
    }
    V v[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; k++) {
        const size_t t = t0 + (size_t) k * blockDim.x;
        if (t < total) v[k] = base[src[k]];
    }
#pragma unroll
    for (int k = 0; k < UNROLL; k++) {
        const size_t t = t0 + (size_t) k * blockDim.x;
        if (t < total) out[t] = v[k];
    }
}

template <typename V>
static int launch_gather_records(const void * base, const u32 * idx, void * out, size_t n,
                                 size_t elsize, cudaStream_t stream)
{
    constexpr int UNROLL = 4;
    const u32 lpr = (u32) (elsize / sizeof(V));
    const size_t total = n * (size_t) lpr;
    const size_t per_block = 256 * UNROLL;
    const size_t blocks = (total + per_block - 1) / per_block;
    if (blocks > 0x7fffffffULL) return (int) cudaErrorInvalidValue;
    gather_records_kernel<V, UNROLL><<<(unsigned) blocks, 256, 0, stream>>>(
        (const V *) base, idx, (V *) out, n, lpr);
    CUDA_LAUNCH_CHECK();
    return 0;
}

extern "C" int mpsk_gather_records(const void * base, const uint32_t * idx, void * out,
        size_t n, size_t elsize, mpsk_stream_t stream_)
{
    if (n == 0 || elsize == 0) return 0;
    cudaStream_t stream = (cudaStream_t) stream_;
    const uintptr_t a = ((uintptr_t) base) | ((uintptr_t) out) | (uintptr_t) elsize;
    if ((a & 15) == 0) return launch_gather_records<uint4>(base, idx, out, n, elsize, stream);
    if ((a & 7) == 0) return launch_gather_records<u64>(base, idx, out, n, elsize, stream);
    if ((a & 3) == 0) return launch_gather_records<u32>(base, idx, out, n, elsize, stream);
    if ((a & 1) == 0) return launch_gather_records<unsigned short>(base, idx, out, n, elsize, stream);
    return launch_gather_records<unsigned char>(base, idx, out, n, elsize, stream);
}

/* ========================================================================= */
/* K4: splitter kernels                                                      */
/* ========================================================================= */

#define MPSK_MAX_KEY_WORDS 16

/* compare key i of the sorted keys (seen through a key view) with cand[]: -1, 0, +1 */
__device__ __forceinline__ int cmp_key(const mpsk_keyview & v, size_t i, const u64 * cand, u32 nw)
{
    const unsigned char * p = (const unsigned char *) v.base + i * v.item_stride;
    for (int w = (int) nw - 1; w >= 0; w--) {
        const u64 k = ((*(const u64 *) (p + (size_t) w * v.word_stride)) ^ v.flip) + (w == 0 ? v.add : 0ULL);
        if (k < cand[w]) return -1;
        if (k > cand[w]) return 1;
    }
    return 0;
}

/* number of keys <= cand (UPPER) or < cand (!UPPER) */
template <bool UPPER>
__device__ __forceinline__ u64 bound_key(const mpsk_keyview & v, size_t n, const u64 * cand, u32 nw)
{
    size_t lo = 0, hi = n;
    while (lo < hi) {
        const size_t mid = lo + ((hi - lo) >> 1);
        const int c = cmp_key(v, mid, cand, nw);
        const bool go_right = UPPER ? (c <= 0) : (c < 0);
        if (go_right) lo = mid + 1; else hi = mid;
    }
    return (u64) lo;
}

__global__ void __launch_bounds__(256)
splitter_count_kernel(mpsk_keyview v, size_t n, u32 nw,
                      const u64 * __restrict__ prefix, int level, u64 * __restrict__ counts)
{
    const u32 b = blockIdx.x;
    const u32 d = threadIdx.x;
    const u32 byteidx = 8 * nw - 1 - (u32) level;   /* from the least significant byte */
    const u32 wi = byteidx >> 3;
    const u32 sh = (byteidx & 7) * 8;
    u64 cand[MPSK_MAX_KEY_WORDS];
    for (u32 w = 0; w < nw; w++) {
        u64 x = prefix[(size_t) b * nw + w];
        if (w < wi) x = ~0ULL;
        else if (w == wi) x |= ((u64) d << sh) | ((sh == 0) ? 0ULL : ((1ULL << sh) - 1ULL));
        cand[w] = x;
    }
    counts[(size_t) b * 256 + d] = bound_key<true>(v, n, cand, nw);
}

extern "C" int mpsk_splitter_count(struct mpsk_keyview view, size_t n, uint32_t nw,
        const uint64_t * prefix, int nsplit, int level, uint64_t * counts, mpsk_stream_t stream)
{
    if (nsplit <= 0) return 0;
    if (nw > MPSK_MAX_KEY_WORDS) return (int) cudaErrorInvalidValue;
    splitter_count_kernel<<<nsplit, 256, 0, (cudaStream_t) stream>>>(
        view, n, nw, (const u64 *) prefix, level, (u64 *) counts);
    CUDA_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(256)
splitter_select_kernel(const u64 * __restrict__ counts, const u64 * __restrict__ target,
                       u64 * __restrict__ prefix, u32 nw, int level)
{
    __shared__ u32 s_min;
    const u32 b = blockIdx.x;
    const u32 d = threadIdx.x;
    if (d == 0) s_min = 255u;
    __syncthreads();
    const bool ok = counts[(size_t) b * 256 + d] >= target[b];
    if (ok) atomicMin(&s_min, d);
    __syncthreads();
    if (d == 0) {
        const u32 byteidx = 8 * nw - 1 - (u32) level;
        const u32 wi = byteidx >> 3;
        const u32 sh = (byteidx & 7) * 8;
        prefix[(size_t) b * nw + wi] |= ((u64) s_min) << sh;
    }
}

extern "C" int mpsk_splitter_select(const uint64_t * counts, const uint64_t * target,
        uint64_t * prefix, uint32_t nw, int nsplit, int level, mpsk_stream_t stream)
{
    if (nsplit <= 0) return 0;
    splitter_select_kernel<<<nsplit, 256, 0, (cudaStream_t) stream>>>(
        (const u64 *) counts, (const u64 *) target, (u64 *) prefix, nw, level);
    CUDA_LAUNCH_CHECK();
    return 0;
}

__global__ void splitter_final_kernel(mpsk_keyview v, size_t n, u32 nw,
                                      const u64 * __restrict__ prefix, int nsplit, u64 * __restrict__ out)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2u * (u32) nsplit) return;
    const u32 b = t % (u32) nsplit;
    const bool upper = t >= (u32) nsplit;
    u64 cand[MPSK_MAX_KEY_WORDS];
    for (u32 w = 0; w < nw; w++) cand[w] = prefix[(size_t) b * nw + w];
    out[t] = upper ? bound_key<true>(v, n, cand, nw) : bound_key<false>(v, n, cand, nw);
}

extern "C" int mpsk_splitter_final(struct mpsk_keyview view, size_t n, uint32_t nw,
        const uint64_t * prefix, int nsplit, uint64_t * out, mpsk_stream_t stream)
{
    if (nsplit <= 0) return 0;
    if (nw > MPSK_MAX_KEY_WORDS) return (int) cudaErrorInvalidValue;
    const int threads = 64;
    const int blocks = (2 * nsplit + threads - 1) / threads;
    splitter_final_kernel<<<blocks, threads, 0, (cudaStream_t) stream>>>(
        view, n, nw, (const u64 *) prefix, nsplit, (u64 *) out);
    CUDA_LAUNCH_CHECK();
    return 0;
}

/*
 * CANDIDATE, off by default (MPSORT_PEER_SPLITTER=1) and not yet run on a GPU: the whole byte-wise
 * descent in ONE kernel per GPU, the per-level all-reduce done over peer memory instead of one
 * ncclAllReduce + two launches per level (8 x ~60 us at 8 GPUs).
 *
 * Every rank owns a mailbox in device memory that all peers have mapped (CUDA IPC; plain pointers
 * for rank threads of one process): counts[2][PEER_MAXS][256] u64 and one flag word per splitter.
 * Block b works on splitter b on every rank. Per level: count the 256 candidates locally
 * (splitter_count_kernel's arithmetic), store them in the own mailbox (parity = level & 1), fence,
 * release-store flag[b] = seq + level + 1; poll the same flag of every peer (acquire), then add the
 * peers' 256 counts read over NVLink and pick the digit (splitter_select_kernel's rule). All ranks
 * compute the same sums, so nothing is broadcast. Two parities suffice: a rank reaches level L+2 only
 * after every peer published level L+1, which a peer does after it has read level L.
 * Block b only ever waits for block b of the peers' kernels; <= 63 blocks are always co-resident. A
 * wait that exceeds `timeout` clock cycles sets *err and leaves (the host aborts the job) instead of
 * hanging the GPU.
 */
#define MPSK_PEER_MAXS 63
struct PeerBoxes { unsigned long long * box[64]; };
__host__ __device__ constexpr size_t peer_box_count_words() { return (size_t) 2 * MPSK_PEER_MAXS * 256; }

__device__ __forceinline__ u32 ld_acquire_sys_u32(const u32 * p)
{
    u32 v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(u32 * p, u32 v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ u64 ld_relaxed_sys_u64(const u64 * p)
{
    u64 v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256)
splitter_descent_peer_kernel(mpsk_keyview v, size_t n, u32 nw, u64 * __restrict__ prefix, const u64 * __restrict__ target,
                             int level0, int nlevels, u32 me, u32 p, PeerBoxes boxes, u32 seq,
                             long long timeout, u32 * __restrict__ err)
{
    __shared__ u64 s_prefix[MPSK_MAX_KEY_WORDS];
    __shared__ u32 s_min, s_abort;
    const u32 b = blockIdx.x, d = threadIdx.x;
    if (d < nw) s_prefix[d] = prefix[(size_t) b * nw + d];
    if (d == 0) s_abort = 0;
    __syncthreads();
    const u64 tgt = target[b];
    u64 * mycounts = boxes.box[me];
    u32 * myflags = (u32 *) (boxes.box[me] + peer_box_count_words());
    for (int level = level0; level < nlevels; level++) {
        const u32 par = (u32) level & 1u;
        const u32 byteidx = 8 * nw - 1 - (u32) level;   /* from the least significant byte */
        const u32 wi = byteidx >> 3;
        const u32 sh = (byteidx & 7) * 8;
        u64 cand[MPSK_MAX_KEY_WORDS];
        for (u32 w = 0; w < nw; w++) {
            u64 x = s_prefix[w];
            if (w < wi) x = ~0ULL;
            else if (w == wi) x |= ((u64) d << sh) | ((sh == 0) ? 0ULL : ((1ULL << sh) - 1ULL));
            cand[w] = x;
        }
        const u64 c = bound_key<true>(v, n, cand, nw);
        const size_t slot = ((size_t) par * MPSK_PEER_MAXS + b) * 256 + d;
        mycounts[slot] = c;
        __threadfence_system();
        if (d == 0) s_min = 255u;
        __syncthreads();
        const u32 want = seq + (u32) level + 1u;
        if (d == 0) st_release_sys_u32(&myflags[b], want);
        if (d < p && d != me) {
            const u32 * pf = (const u32 *) (boxes.box[d] + peer_box_count_words()) + b;
            const long long t0 = clock64();
            while ((int) (ld_acquire_sys_u32(pf) - want) < 0) {
                if (clock64() - t0 > timeout) { s_abort = 1; break; }
                __nanosleep(100);
            }
        }
        __syncthreads();
        if (s_abort) {
            if (d == 0) atomicExch(err, 1u);
            return;
        }
        __threadfence_system();
        u64 sum = c;
        for (u32 r = 0; r < p; r++)
            if (r != me) sum += ld_relaxed_sys_u64(boxes.box[r] + slot);
        if (sum >= tgt) atomicMin(&s_min, d);
        __syncthreads();
        if (d == 0) s_prefix[wi] |= ((u64) s_min) << sh;
        __syncthreads();
    }
    if (d < nw) prefix[(size_t) b * nw + d] = s_prefix[d];
}

extern "C" size_t mpsk_peer_box_bytes(void) { return peer_box_count_words() * sizeof(u64) + 256 * sizeof(u32); }

extern "C" int mpsk_splitter_descent_peer(struct mpsk_keyview view, size_t n, uint32_t nw,
        uint64_t * prefix, const uint64_t * target, int nsplit, int level0, int nlevels,
        uint32_t me, uint32_t p, void * const * boxes, uint32_t seq, uint32_t * err, mpsk_stream_t stream)
{
    if (nsplit <= 0 || level0 >= nlevels) return 0;
    if (nw > MPSK_MAX_KEY_WORDS || nsplit > MPSK_PEER_MAXS || p > 64 || me >= p) return (int) cudaErrorInvalidValue;
    PeerBoxes pb;
    for (u32 r = 0; r < 64; r++) pb.box[r] = r < p ? (unsigned long long *) boxes[r] : NULL;
    /* ~10 s at 2 GHz: a peer may still be in its local sort; a dead peer must not hang the box */
    const long long timeout = 20000000000LL;
    splitter_descent_peer_kernel<<<nsplit, 256, 0, (cudaStream_t) stream>>>(
        view, n, nw, (u64 *) prefix, (const u64 *) target, level0, nlevels, me, p, pb, seq, timeout, err);
    CUDA_LAUNCH_CHECK();
    return 0;
}

#define MPSK_MAX_SUM_SRCS 64
struct SumSrcs { const u64 * p[MPSK_MAX_SUM_SRCS]; };

__global__ void sum_u64_kernel(u64 * __restrict__ dst, SumSrcs srcs, int nsrc, size_t count)
{
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    u64 s = 0;
    for (int k = 0; k < nsrc; k++) s += srcs.p[k][i];
    dst[i] = s;
}

extern "C" int mpsk_sum_u64(uint64_t * dst, const uint64_t * const * srcs, int nsrc, size_t count,
        mpsk_stream_t stream)
{
    if (count == 0) return 0;
    if (nsrc > MPSK_MAX_SUM_SRCS) return (int) cudaErrorInvalidValue;
    SumSrcs s;
    for (int k = 0; k < nsrc; k++) s.p[k] = (const u64 *) srcs[k];
    const int threads = 256;
    const size_t blocks = (count + threads - 1) / threads;
    sum_u64_kernel<<<(unsigned) blocks, threads, 0, (cudaStream_t) stream>>>((u64 *) dst, s, nsrc, count);
    CUDA_LAUNCH_CHECK();
    return 0;
}

/* ========================================================================= */
/* K7: stable p-way merge of the received runs (replaces the second radix_sort, */
/* mpsort-mpi.c:597, whose input is p sorted runs in source-rank order)          */
/* ========================================================================= */
/*
 * The receive buffer holds p sorted runs (run r = records [rdispl[r], rdispl[r+1])).
 * 1. merge_sample_kernel: every S-th key of every run (the last key of each full
 *    block of S) -> samples in (run, position) order.
 * 2. the samples are sorted stably by key with the onesweep sort (host side), which
 *    orders them by (key, run, position).
 * 3. merge_bounds_kernel: every k-th merged sample is a tile boundary; its cut
 *    position in every run is found by binary search (upper bound in lower runs,
 *    lower bound in higher runs: ties go to the lower run first, like the stable
 *    merge of stdlib/msort.c:78). A tile holds < (k + p) * S records.
 * 4. merge_tile_kernel: one CTA per tile loads the p sub-ranges' keys into shared
 *    memory, merges them pairwise in log2(p) rounds (every key finds its rank in the
 *    sibling sequence by binary search: A-side lower bound, B-side upper bound) and
 *    writes the records out in merged order.
 * HBM traffic: E read + E write per record (+ ~8/S for the samples).
 */
#define MPSK_MERGE_MAX_RUNS 32
#define MPSK_MERGE_TILE 4096
#ifndef MPSK_MERGE_THREADS
#define MPSK_MERGE_THREADS 512
#endif

struct MergeRuns {
    u32 p;
    u32 S;            /* sample stride */
    u32 k;            /* samples per tile */
    u32 rdispl[MPSK_MERGE_MAX_RUNS + 1];   /* run starts in records */
    u32 sstart[MPSK_MERGE_MAX_RUNS + 1];   /* first sample id of every run */
};

__device__ __forceinline__ u64 load_key_any(const unsigned char * rec, const KeyDesc & d, bool fast8)
{
    if (fast8) return (*(const u64 *) (rec + d.offset)) ^ (d.is_signed ? (1ULL << 63) : 0ULL);
    return pack_key_word(rec, d);
}

__global__ void __launch_bounds__(256)
merge_sample_kernel(const unsigned char * __restrict__ recv, KeyDesc d, bool fast8, MergeRuns m,
                    u64 * __restrict__ skeys)
{
    const u32 ns = m.sstart[m.p];
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < ns; s += gridDim.x * blockDim.x) {
        u32 r = 0;
        while (s >= m.sstart[r + 1]) r++;
        const u32 j = s - m.sstart[r];
        const size_t pos = (size_t) m.rdispl[r] + (size_t) (j + 1) * m.S - 1;
        skeys[s] = load_key_any(recv + pos * d.elsize, d, fast8);
    }
}

/* cut[t * p + r] for t = 0 .. ntiles */
__global__ void __launch_bounds__(256)
merge_bounds_kernel(const unsigned char * __restrict__ recv, KeyDesc d, bool fast8, MergeRuns m,
                    const u64 * __restrict__ sorted_skeys, const u32 * __restrict__ sorted_sid,
                    u32 ntiles, u32 * __restrict__ cut)
{
    const u32 total = (ntiles + 1) * m.p;
    for (u32 x = blockIdx.x * blockDim.x + threadIdx.x; x < total; x += gridDim.x * blockDim.x) {
        const u32 t = x / m.p, r = x - t * m.p;
        const u32 len = m.rdispl[r + 1] - m.rdispl[r];
        u32 c;
        if (t == 0) c = 0;
        else if (t == ntiles) c = len;
        else {
            const u32 mi = t * m.k - 1;
            const u64 kb = sorted_skeys[mi];
            const u32 sid = sorted_sid[mi];
            u32 rb = 0;
            while (sid >= m.sstart[rb + 1]) rb++;
            if (r == rb) {
                c = (sid - m.sstart[rb] + 1) * m.S;
            } else {
                const unsigned char * base = recv + (size_t) m.rdispl[r] * d.elsize;
                u32 lo = 0, hi = len;
                const bool upper = r < rb;
                while (lo < hi) {
                    const u32 mid = lo + ((hi - lo) >> 1);
                    const u64 kk = load_key_any(base + (size_t) mid * d.elsize, d, fast8);
                    const bool right = upper ? (kk <= kb) : (kk < kb);
                    if (right) lo = mid + 1; else hi = mid;
                }
                c = lo;
            }
        }
        cut[x] = c;
    }
}

/* shared-memory index with one pad slot per 8 items: a thread's 8 consecutive outputs
 * are 64 bytes apart from its neighbour's, which would be a 16-way bank conflict */
#define MPD(i) ((i) + ((i) >> 3))
#define MPSK_MERGE_PADDED (MPSK_MERGE_TILE + MPSK_MERGE_TILE / 8)

/* FAST8: one aligned 8-byte key word (no generic key packing code in the kernel);
 * LPR1: a record is exactly one V (no division in the output loop) */
template <typename V, bool FAST8, bool LPR1>
__global__ void __launch_bounds__(MPSK_MERGE_THREADS, 2)
merge_tile_kernel(const unsigned char * __restrict__ recv, KeyDesc d, MergeRuns m,
                  const u32 * __restrict__ cut, unsigned char * __restrict__ out, u32 * __restrict__ overflow)
{
    constexpr int VT = MPSK_MERGE_TILE / MPSK_MERGE_THREADS;      /* items per thread */
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 * kA = (u64 *) smem_raw;
    u64 * kB = kA + MPSK_MERGE_PADDED;
    u32 * sA = (u32 *) (kB + MPSK_MERGE_PADDED);
    u32 * sB = sA + MPSK_MERGE_PADDED;
    __shared__ u32 seqoff[MPSK_MERGE_MAX_RUNS + 1];
    __shared__ u32 srcbase[MPSK_MERGE_MAX_RUNS];
    __shared__ u32 s_outstart;

    const u32 t = blockIdx.x, p = m.p, tid = threadIdx.x;
    if (tid < 32) {
        /* one lane per run (p <= 32): the 2p cut words are fetched in parallel, not by one
         * thread in a dependent loop (that loop alone was ~4 us per tile at p = 8) */
        u32 c0 = 0, c1 = 0;
        if (tid < p) { c0 = cut[t * p + tid]; c1 = cut[(t + 1) * p + tid]; }
        const u32 len = c1 - c0;
        u32 incl = len, sum0 = c0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 y = __shfl_up_sync(FULL_MASK, incl, o);
            if (tid >= (u32) o) incl += y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum0 += __shfl_xor_sync(FULL_MASK, sum0, o);
        if (tid < p) { seqoff[tid] = incl - len; srcbase[tid] = m.rdispl[tid] + c0; }
        if (tid == p - 1) seqoff[p] = incl;
        if (tid == 0) s_outstart = sum0;
    }
    __syncthreads();
    const u32 cnt = seqoff[p];
    if (cnt > MPSK_MERGE_TILE) {            /* cannot happen (tile bound); never corrupt memory */
        if (tid == 0) atomicAdd(overflow, 1u);
        return;
    }
    /* ---- load the keys of the p sub-ranges, run-major; all loads of a thread in flight */
    {
        u32 src[VT];
        u64 key[VT];
        u32 r = 0;                                  /* a thread's positions increase: the run index only moves up */
#pragma unroll
        for (int k = 0; k < VT; k++) {
            const u32 i = tid + k * MPSK_MERGE_THREADS;
            src[k] = 0;
            if (i < cnt) {
                while (i >= seqoff[r + 1]) r++;
                src[k] = srcbase[r] + (i - seqoff[r]);
            }
        }
#pragma unroll
        for (int k = 0; k < VT; k++) {
            const u32 i = tid + k * MPSK_MERGE_THREADS;
            if (i < cnt) key[k] = load_key_any(recv + (size_t) src[k] * d.elsize, d, FAST8);
        }
#pragma unroll
        for (int k = 0; k < VT; k++) {
            const u32 i = tid + k * MPSK_MERGE_THREADS;
            if (i < cnt) { kA[MPD(i)] = key[k]; sA[MPD(i)] = src[k]; }
        }
    }
    __syncthreads();
    /* ---- pairwise merge rounds over groups of w runs: every thread produces VT
     * consecutive outputs, starting from its merge-path intersection (one binary
     * search per thread and pair instead of one per item). Ties take from A, the
     * lower runs: stable. */
    for (u32 w = 1; w < p; w <<= 1) {
        u32 o = tid * VT;
        const u32 end = min(o + (u32) VT, cnt);
        u32 g = 0;                                  /* pair index: groups 2g and 2g+1 */
        while (o < end) {
            while (seqoff[min((2 * g + 2) * w, p)] <= o) g++;
            const u32 a0 = seqoff[min(2 * g * w, p)];
            const u32 a1 = seqoff[min((2 * g + 1) * w, p)];
            const u32 b1 = seqoff[min((2 * g + 2) * w, p)];
            const u32 lenA = a1 - a0, lenB = b1 - a1;
            const u32 seg_end = min(end, b1);
            const u32 diag = o - a0;
            u32 lo = diag > lenB ? diag - lenB : 0, hi = min(diag, lenA);
            while (lo < hi) {
                const u32 mid = (lo + hi) >> 1;
                if (kA[MPD(a0 + mid)] <= kA[MPD(a1 + diag - 1 - mid)]) lo = mid + 1; else hi = mid;
            }
            /* branch-free sequential merge: both candidates (key, source) live in registers,
             * the one taken is replaced by its successor (index clamped at the sequence end) */
            u32 ai = a0 + lo, bi = a1 + (diag - lo);            /* absolute positions */
            u64 ka = kA[MPD(min(ai, b1 - 1))], kb = kA[MPD(min(bi, b1 - 1))];
            u32 sa = sA[MPD(min(ai, b1 - 1))], sb = sA[MPD(min(bi, b1 - 1))];
            for (; o < seg_end; o++) {
                const bool takeA = (bi >= b1) || (ai < a1 && ka <= kb);
                kB[MPD(o)] = takeA ? ka : kb;
                sB[MPD(o)] = takeA ? sa : sb;
                ai += takeA ? 1u : 0u;
                bi += takeA ? 0u : 1u;
                const u32 nxt = min(takeA ? ai : bi, b1 - 1);
                const u64 nk = kA[MPD(nxt)];
                const u32 ns = sA[MPD(nxt)];
                ka = takeA ? nk : ka; sa = takeA ? ns : sa;
                kb = takeA ? kb : nk; sb = takeA ? sb : ns;
            }
        }
        __syncthreads();
        u64 * tk = kA; kA = kB; kB = tk;
        u32 * ts = sA; sA = sB; sB = ts;
    }
    /* ---- write the records in merged order (lanes of one record move consecutive pieces) */
    const u32 lpr = LPR1 ? 1u : (u32) (d.elsize / sizeof(V));
    const V * in = (const V *) recv;
    V * o = (V *) out + (size_t) s_outstart * lpr;
    const u32 totalv = cnt * lpr;
    for (u32 x0 = 0; x0 < totalv; x0 += VT * MPSK_MERGE_THREADS) {
        V v[VT];
#pragma unroll
        for (int k = 0; k < VT; k++) {
            const u32 x = x0 + tid + k * MPSK_MERGE_THREADS;
            if (x < totalv) {
                if (LPR1) {
                    v[k] = in[sA[MPD(x)]];
                } else {
                    const u32 i = x / lpr, part = x - i * lpr;
                    v[k] = in[(size_t) sA[MPD(i)] * lpr + part];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < VT; k++) {
            const u32 x = x0 + tid + k * MPSK_MERGE_THREADS;
            if (x < totalv) o[x] = v[k];
        }
    }
}

/* ---- the same merge for 16-byte records {u64 key, u64 other}: the records themselves
 * are staged and merged in shared memory, so every record is read once and written
 * once with coalesced 16-byte accesses (no key pre-read, no gather by index). */
#define MPSK_MERGE16_TILE 2048
#define MPSK_MERGE16_THREADS 512
#define MPSK_MERGE16_PADDED (MPSK_MERGE16_TILE + MPSK_MERGE16_TILE / 8)

template <bool KHI>
__global__ void __launch_bounds__(MPSK_MERGE16_THREADS, 3)
merge_tile_rec16_kernel(const uint4 * __restrict__ recv, u64 flip, MergeRuns m,
                        const u32 * __restrict__ cut, uint4 * __restrict__ out, u32 * __restrict__ overflow)
{
    constexpr int VT = MPSK_MERGE16_TILE / MPSK_MERGE16_THREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 * A = (uint4 *) smem_raw;
    uint4 * B = A + MPSK_MERGE16_PADDED;
    __shared__ u32 seqoff[MPSK_MERGE_MAX_RUNS + 1];
    __shared__ u32 srcbase[MPSK_MERGE_MAX_RUNS];
    __shared__ u32 s_outstart;

    const u32 t = blockIdx.x, p = m.p, tid = threadIdx.x;
    if (tid < 32) {
        /* one lane per run (p <= 32): the 2p cut words are fetched in parallel, not by one
         * thread in a dependent loop (that loop alone was ~4 us per tile at p = 8) */
        u32 c0 = 0, c1 = 0;
        if (tid < p) { c0 = cut[t * p + tid]; c1 = cut[(t + 1) * p + tid]; }
        const u32 len = c1 - c0;
        u32 incl = len, sum0 = c0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 y = __shfl_up_sync(FULL_MASK, incl, o);
            if (tid >= (u32) o) incl += y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum0 += __shfl_xor_sync(FULL_MASK, sum0, o);
        if (tid < p) { seqoff[tid] = incl - len; srcbase[tid] = m.rdispl[tid] + c0; }
        if (tid == p - 1) seqoff[p] = incl;
        if (tid == 0) s_outstart = sum0;
    }
    __syncthreads();
    const u32 cnt = seqoff[p];
    if (cnt > MPSK_MERGE16_TILE) {
        if (tid == 0) atomicAdd(overflow, 1u);
        return;
    }
    {
        uint4 rec[VT];
#pragma unroll
        for (int k = 0; k < VT; k++) {
            const u32 i = tid + k * MPSK_MERGE16_THREADS;
            if (i < cnt) {
                u32 r = 0;
                while (i >= seqoff[r + 1]) r++;
                rec[k] = recv[srcbase[r] + (i - seqoff[r])];
            }
        }
#pragma unroll
        for (int k = 0; k < VT; k++) {
            const u32 i = tid + k * MPSK_MERGE16_THREADS;
            if (i < cnt) A[MPD(i)] = rec[k];
        }
    }
    __syncthreads();
#define KEY16(arr, i) (((const u64 *) &(arr)[MPD(i)])[KHI ? 1 : 0] ^ flip)
    for (u32 w = 1; w < p; w <<= 1) {
        u32 o = tid * VT;
        const u32 end = min(o + (u32) VT, cnt);
        u32 g = 0;
        while (o < end) {
            while (seqoff[min((2 * g + 2) * w, p)] <= o) g++;
            const u32 a0 = seqoff[min(2 * g * w, p)];
            const u32 a1 = seqoff[min((2 * g + 1) * w, p)];
            const u32 b1 = seqoff[min((2 * g + 2) * w, p)];
            const u32 lenA = a1 - a0, lenB = b1 - a1;
            const u32 seg_end = min(end, b1);
            const u32 diag = o - a0;
            u32 lo = diag > lenB ? diag - lenB : 0, hi = min(diag, lenA);
            while (lo < hi) {
                const u32 mid = (lo + hi) >> 1;
                if (KEY16(A, a0 + mid) <= KEY16(A, a1 + diag - 1 - mid)) lo = mid + 1; else hi = mid;
            }
            u32 ai = lo, bi = diag - lo;
            u64 ka = ai < lenA ? KEY16(A, a0 + ai) : 0, kb = bi < lenB ? KEY16(A, a1 + bi) : 0;
            for (; o < seg_end; o++) {
                const bool takeA = (bi >= lenB) || (ai < lenA && ka <= kb);
                if (takeA) {
                    B[MPD(o)] = A[MPD(a0 + ai)];
                    ai++;
                    if (ai < lenA) ka = KEY16(A, a0 + ai);
                } else {
                    B[MPD(o)] = A[MPD(a1 + bi)];
                    bi++;
                    if (bi < lenB) kb = KEY16(A, a1 + bi);
                }
            }
        }
        __syncthreads();
        uint4 * tmp = A; A = B; B = tmp;
    }
#undef KEY16
    uint4 * o = out + s_outstart;
#pragma unroll
    for (int k = 0; k < VT; k++) {
        const u32 i = tid + k * MPSK_MERGE16_THREADS;
        if (i < cnt) o[i] = A[MPD(i)];
    }
}

static bool merge_rec16_ok(const void * recv, const void * out, size_t elsize, size_t offset, uint32_t width, uint32_t nwords)
{
    static int disabled = -1;
    if (disabled < 0) disabled = getenv("MPSORT_NO_MERGE16") ? 1 : 0;
    if (disabled) return false;
    return elsize == 16 && width == 8 && nwords == 1 && (offset == 0 || offset == 8)
           && ((((uintptr_t) recv) | ((uintptr_t) out)) & 15) == 0;
}

/* the record-staging kernel wins for two runs (one merge round); with more rounds the
 * (key, index) kernel moves fewer bytes per round (profiles/r01_merge_kernels.log) */
extern "C" size_t mpsk_merge_tile_items_for(const void * recv, const void * out, size_t elsize, size_t offset,
        uint32_t width, uint32_t nwords, uint32_t p)
{
    return (p == 2 && merge_rec16_ok(recv, out, elsize, offset, width, nwords)) ? MPSK_MERGE16_TILE : MPSK_MERGE_TILE;
}

extern "C" size_t mpsk_merge_tile_items(void) { return MPSK_MERGE_TILE; }

extern "C" int mpsk_merge_samples(const void * recv, size_t elsize, size_t offset, uint32_t width,
        uint32_t nwords, int is_signed, uint32_t p, uint32_t S, uint32_t k,
        const uint32_t * rdispl, const uint32_t * sstart, uint64_t * skeys, mpsk_stream_t stream)
{
    if (p > MPSK_MERGE_MAX_RUNS) return (int) cudaErrorInvalidValue;
    MergeRuns m; m.p = p; m.S = S; m.k = k;
    for (u32 r = 0; r <= p; r++) { m.rdispl[r] = rdispl[r]; m.sstart[r] = sstart[r]; }
    const u32 ns = sstart[p];
    if (ns == 0) return 0;
    KeyDesc d; d.elsize = elsize; d.offset = offset; d.width = width; d.nwords = nwords; d.is_signed = is_signed; d.g = 0; d.sub = 0;
    const bool fast8 = (width == 8) && (nwords == 1) && (elsize % 8 == 0) && (offset % 8 == 0) && ((((uintptr_t) recv) & 7) == 0);
    u32 blocks = (ns + 255) / 256;
    if (blocks > (u32) num_sms() * 8) blocks = (u32) num_sms() * 8;
    merge_sample_kernel<<<blocks, 256, 0, (cudaStream_t) stream>>>((const unsigned char *) recv, d, fast8, m, (u64 *) skeys);
    CUDA_LAUNCH_CHECK();
    return 0;
}

template <typename V>
static int launch_merge_tiles(const void * recv, KeyDesc d, bool fast8, const MergeRuns & m, const u32 * cut,
                              void * out, u32 * overflow, u32 ntiles, cudaStream_t stream)
{
    const int smem = MPSK_MERGE_PADDED * (8 + 8 + 4 + 4);
    const bool lpr1 = d.elsize == sizeof(V);
#define MERGE_LAUNCH(F8, L1) do { \
        auto kern = merge_tile_kernel<V, F8, L1>; \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
        if (e != cudaSuccess) return (int) e; \
        kern<<<ntiles, MPSK_MERGE_THREADS, smem, stream>>>((const unsigned char *) recv, d, m, cut, \
                                                           (unsigned char *) out, overflow); } while (0)
    if (fast8 && lpr1) MERGE_LAUNCH(true, true);
    else if (fast8) MERGE_LAUNCH(true, false);
    else if (lpr1) MERGE_LAUNCH(false, true);
    else MERGE_LAUNCH(false, false);
#undef MERGE_LAUNCH
    CUDA_LAUNCH_CHECK();
    return 0;
}

/*
 * CANDIDATE, off by default (MPSORT_MERGE_BUCKET=1 selects it) and not yet run on a GPU: the tile
 * merge in ONE round by interpolation buckets instead of log2(p) merge-path rounds; the idea, its
 * exactness argument and the arithmetic are in mpsort_merge_bucket.cuh (that part is checked on the
 * CPU by tests/test_merge_bucket_emul.py). Same grid, threads, shared-memory size and tile bounds
 * as merge_tile_kernel. Keys and source positions stay in registers from the load to the scatter;
 * a tile that fails the spread test (bucket longer than CMAX, key outside the boundary keys, first
 * or last tile of a part whose range is open) stages them like merge_tile_kernel and runs that
 * kernel's rounds. overflow[1] counts those tiles.
 */
template <typename V, bool FAST8, bool LPR1>
__global__ void __launch_bounds__(MPSK_MERGE_THREADS, 2)
merge_tile_bucket_kernel(const unsigned char * __restrict__ recv, KeyDesc d, MergeRuns m,
                         const u32 * __restrict__ cut, const u64 * __restrict__ bkeys, u32 ntiles,
                         unsigned char * __restrict__ out, u32 * __restrict__ overflow)
{
    constexpr int VT = MPSK_MERGE_TILE / MPSK_MERGE_THREADS;      /* items per thread */
    static_assert(mbk::NB / 16 == MPSK_MERGE_THREADS, "one thread scans 16 consecutive bucket counters");
    static_assert(VT <= 8 && mbk::CMAX <= 16, "arrival slots are packed as nibbles of one word");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    /* bucket path */
    u32 * cntp = (u32 *) smem_raw;                                  /* [NBP] counters, later bucket starts */
    u64 * skey = (u64 *) (smem_raw + mbk::NBP * 4);                 /* [TILE] keys in bucket order */
    u32 * ssrc = (u32 *) (skey + MPSK_MERGE_TILE);                  /* [TILE] their source positions */
    u32 * osrc = ssrc + MPSK_MERGE_TILE;                            /* [TILE] source positions in merged order */
    static_assert((mbk::NBP * 4) % 16 == 0, "skey must stay 8-byte aligned");
    static_assert(mbk::NBP * 4 + MPSK_MERGE_TILE * (8 + 4 + 4) <= MPSK_MERGE_PADDED * (8 + 8 + 4 + 4), "fits the merge kernel's shared memory");
    /* fallback: the layout of merge_tile_kernel (aliases the above, which is dead by then) */
    u64 * kA = (u64 *) smem_raw;
    u64 * kB = kA + MPSK_MERGE_PADDED;
    u32 * sA = (u32 *) (kB + MPSK_MERGE_PADDED);
    u32 * sB = sA + MPSK_MERGE_PADDED;
    __shared__ u32 seqoff[MPSK_MERGE_MAX_RUNS + 1];
    __shared__ u32 srcbase[MPSK_MERGE_MAX_RUNS];
    __shared__ u32 s_outstart;
    __shared__ u64 s_bound[2];
    __shared__ u32 s_wsum[MPSK_MERGE_THREADS / 32];

    const u32 t = blockIdx.x, p = m.p, tid = threadIdx.x;
    if (tid < 32) {
        u32 c0 = 0, c1 = 0;
        if (tid < p) { c0 = cut[t * p + tid]; c1 = cut[(t + 1) * p + tid]; }
        const u32 len = c1 - c0;
        u32 incl = len, sum0 = c0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 y = __shfl_up_sync(FULL_MASK, incl, o);
            if (tid >= (u32) o) incl += y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum0 += __shfl_xor_sync(FULL_MASK, sum0, o);
        if (tid < p) { seqoff[tid] = incl - len; srcbase[tid] = m.rdispl[tid] + c0; }
        if (tid == p - 1) seqoff[p] = incl;
        if (tid == 0) s_outstart = sum0;
    } else if (tid == 32) {
        /* the tile's keys lie between the boundary samples of its two cuts (merge_bounds_kernel);
         * the first and the last tile of a part are open on one side */
        s_bound[0] = t > 0 ? bkeys[(size_t) t * m.k - 1] : 0ULL;
        s_bound[1] = t + 1 < ntiles ? bkeys[(size_t) (t + 1) * m.k - 1] : ~0ULL;
    }
    for (u32 i = tid; i < mbk::NBP; i += MPSK_MERGE_THREADS) cntp[i] = 0;
    __syncthreads();
    const u32 cnt = seqoff[p];
    if (cnt > MPSK_MERGE_TILE) {            /* cannot happen (tile bound); never corrupt memory */
        if (tid == 0) atomicAdd(overflow, 1u);
        return;
    }
    const u64 klo = s_bound[0], khi = s_bound[1];
    const u32 sh = mbk::shift_for(klo, khi);

    /* ---- load the keys of the p sub-ranges, run-major (as merge_tile_kernel); they stay in registers */
    u32 src[VT];
    u64 key[VT];
    {
        u32 r = 0;
#pragma unroll
        for (int k = 0; k < VT; k++) {
            const u32 i = tid + k * MPSK_MERGE_THREADS;
            src[k] = 0;
            if (i < cnt) {
                while (i >= seqoff[r + 1]) r++;
                src[k] = srcbase[r] + (i - seqoff[r]);
            }
        }
#pragma unroll
        for (int k = 0; k < VT; k++) {
            const u32 i = tid + k * MPSK_MERGE_THREADS;
            key[k] = 0;
            if (i < cnt) key[k] = load_key_any(recv + (size_t) src[k] * d.elsize, d, FAST8);
        }
    }
    /* ---- count: arrival slot of every record in its bucket */
    u32 slots = 0;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < VT; k++) {
        const u32 i = tid + k * MPSK_MERGE_THREADS;
        if (i < cnt) {
            const bool ok = mbk::in_range(key[k], klo, khi);
            bad |= !ok;
            const u32 b = ok ? mbk::bucket_of(key[k], klo, sh) : 0u;
            const u32 s = atomicAdd(&cntp[mbk::padc(b < mbk::NB ? b : mbk::NB - 1)], 1u);
            slots |= (s < 15u ? s : 15u) << (4 * k);
        }
    }
    __syncthreads();
    /* ---- exclusive scan of the counters; thread tid owns buckets [16 tid, 16 tid + 16) */
    u32 tot = 0, mx = 0;
    {
        const u32 * mine = cntp + 17 * tid;
#pragma unroll
        for (int j = 0; j < 16; j++) { const u32 c = mine[j]; tot += c; mx = c > mx ? c : mx; }
    }
    u32 incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 y = __shfl_up_sync(FULL_MASK, incl, o);
        if ((tid & 31u) >= (u32) o) incl += y;
    }
    if ((tid & 31u) == 31u) s_wsum[tid >> 5] = incl;
    const int fallback = __syncthreads_or((bad || mx > mbk::CMAX) ? 1 : 0);

    const u32 lpr = LPR1 ? 1u : (u32) (d.elsize / sizeof(V));
    const V * in = (const V *) recv;
    V * o = (V *) out + (size_t) s_outstart * lpr;
    const u32 totalv = cnt * lpr;

    if (fallback) {
        /* ---- not spread enough: the merge-path rounds of merge_tile_kernel */
        if (tid == 0) atomicAdd(overflow + 1, 1u);
#pragma unroll
        for (int k = 0; k < VT; k++) {
            const u32 i = tid + k * MPSK_MERGE_THREADS;
            if (i < cnt) { kA[MPD(i)] = key[k]; sA[MPD(i)] = src[k]; }
        }
        __syncthreads();
        for (u32 w = 1; w < p; w <<= 1) {
            u32 oo = tid * VT;
            const u32 end = min(oo + (u32) VT, cnt);
            u32 g = 0;
            while (oo < end) {
                while (seqoff[min((2 * g + 2) * w, p)] <= oo) g++;
                const u32 a0 = seqoff[min(2 * g * w, p)];
                const u32 a1 = seqoff[min((2 * g + 1) * w, p)];
                const u32 b1 = seqoff[min((2 * g + 2) * w, p)];
                const u32 lenA = a1 - a0, lenB = b1 - a1;
                const u32 seg_end = min(end, b1);
                const u32 diag = oo - a0;
                u32 lo = diag > lenB ? diag - lenB : 0, hi = min(diag, lenA);
                while (lo < hi) {
                    const u32 mid = (lo + hi) >> 1;
                    if (kA[MPD(a0 + mid)] <= kA[MPD(a1 + diag - 1 - mid)]) lo = mid + 1; else hi = mid;
                }
                u32 ai = a0 + lo, bi = a1 + (diag - lo);
                u64 ka = kA[MPD(min(ai, b1 - 1))], kb = kA[MPD(min(bi, b1 - 1))];
                u32 sa = sA[MPD(min(ai, b1 - 1))], sb = sA[MPD(min(bi, b1 - 1))];
                for (; oo < seg_end; oo++) {
                    const bool takeA = (bi >= b1) || (ai < a1 && ka <= kb);
                    kB[MPD(oo)] = takeA ? ka : kb;
                    sB[MPD(oo)] = takeA ? sa : sb;
                    ai += takeA ? 1u : 0u;
                    bi += takeA ? 0u : 1u;
                    const u32 nxt = min(takeA ? ai : bi, b1 - 1);
                    const u64 nk = kA[MPD(nxt)];
                    const u32 ns = sA[MPD(nxt)];
                    ka = takeA ? nk : ka; sa = takeA ? ns : sa;
                    kb = takeA ? kb : nk; sb = takeA ? sb : ns;
                }
            }
            __syncthreads();
            u64 * tk = kA; kA = kB; kB = tk;
            u32 * ts = sA; sA = sB; sB = ts;
        }
        for (u32 x0 = 0; x0 < totalv; x0 += VT * MPSK_MERGE_THREADS) {
            V v[VT];
#pragma unroll
            for (int k = 0; k < VT; k++) {
                const u32 x = x0 + tid + k * MPSK_MERGE_THREADS;
                if (x < totalv) {
                    if (LPR1) {
                        v[k] = in[sA[MPD(x)]];
                    } else {
                        const u32 i = x / lpr, part = x - i * lpr;
                        v[k] = in[(size_t) sA[MPD(i)] * lpr + part];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < VT; k++) {
                const u32 x = x0 + tid + k * MPSK_MERGE_THREADS;
                if (x < totalv) o[x] = v[k];
            }
        }
        return;
    }

    /* ---- bucket starts (second read of the counters keeps the register count down) */
    {
        u32 run = incl - tot;
        for (u32 w = 0; w < (tid >> 5); w++) run += s_wsum[w];
        u32 * mine = cntp + 17 * tid;
#pragma unroll
        for (int j = 0; j < 16; j++) { const u32 c = mine[j]; mine[j] = run; run += c; }
    }
    __syncthreads();
    /* ---- scatter (key, source position) to bucket order */
#pragma unroll
    for (int k = 0; k < VT; k++) {
        const u32 i = tid + k * MPSK_MERGE_THREADS;
        if (i < cnt) {
            const u32 pos = cntp[mbk::padc(mbk::bucket_of(key[k], klo, sh))] + ((slots >> (4 * k)) & 15u);
            skey[pos] = key[k];
            ssrc[pos] = src[k];
        }
    }
    __syncthreads();
    /* ---- rank inside the bucket: the merged order of the source positions */
#pragma unroll
    for (int k = 0; k < VT; k++) {
        const u32 pos = tid + k * MPSK_MERGE_THREADS;
        if (pos < cnt) osrc[mbk::merged_position(cntp, skey, ssrc, pos, cnt, klo, sh)] = ssrc[pos];
    }
    __syncthreads();
    /* ---- write the records in merged order (as merge_tile_kernel) */
    for (u32 x0 = 0; x0 < totalv; x0 += VT * MPSK_MERGE_THREADS) {
        V v[VT];
#pragma unroll
        for (int k = 0; k < VT; k++) {
            const u32 x = x0 + tid + k * MPSK_MERGE_THREADS;
            if (x < totalv) {
                if (LPR1) {
                    v[k] = in[osrc[x]];
                } else {
                    const u32 i = x / lpr, part = x - i * lpr;
                    v[k] = in[(size_t) osrc[i] * lpr + part];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < VT; k++) {
            const u32 x = x0 + tid + k * MPSK_MERGE_THREADS;
            if (x < totalv) o[x] = v[k];
        }
    }
}

template <typename V>
static int launch_merge_bucket_tiles(const void * recv, KeyDesc d, bool fast8, const MergeRuns & m, const u32 * cut,
                                     const u64 * bkeys, void * out, u32 * overflow, u32 ntiles, cudaStream_t stream)
{
    const int smem = MPSK_MERGE_PADDED * (8 + 8 + 4 + 4);
    const bool lpr1 = d.elsize == sizeof(V);
#define MERGE_LAUNCH(F8, L1) do { \
        auto kern = merge_tile_bucket_kernel<V, F8, L1>; \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
        if (e != cudaSuccess) return (int) e; \
        kern<<<ntiles, MPSK_MERGE_THREADS, smem, stream>>>((const unsigned char *) recv, d, m, cut, bkeys, ntiles, \
                                                           (unsigned char *) out, overflow); } while (0)
    if (fast8 && lpr1) MERGE_LAUNCH(true, true);
    else if (fast8) MERGE_LAUNCH(true, false);
    else if (lpr1) MERGE_LAUNCH(false, true);
    else MERGE_LAUNCH(false, false);
#undef MERGE_LAUNCH
    CUDA_LAUNCH_CHECK();
    return 0;
}

extern "C" int mpsk_merge_runs(const void * recv, void * out, size_t elsize, size_t offset, uint32_t width,
        uint32_t nwords, int is_signed, uint32_t p, uint32_t S, uint32_t k,
        const uint32_t * rdispl, const uint32_t * sstart,
        const uint64_t * sorted_skeys, const uint32_t * sorted_sid, uint32_t ntiles,
        uint32_t * cut, uint32_t * overflow, mpsk_stream_t stream_)
{
    if (p > MPSK_MERGE_MAX_RUNS) return (int) cudaErrorInvalidValue;
    cudaStream_t stream = (cudaStream_t) stream_;
    MergeRuns m; m.p = p; m.S = S; m.k = k;
    for (u32 r = 0; r <= p; r++) { m.rdispl[r] = rdispl[r]; m.sstart[r] = sstart[r]; }
    KeyDesc d; d.elsize = elsize; d.offset = offset; d.width = width; d.nwords = nwords; d.is_signed = is_signed; d.g = 0; d.sub = 0;
    const bool fast8 = (width == 8) && (nwords == 1) && (elsize % 8 == 0) && (offset % 8 == 0) && ((((uintptr_t) recv) & 7) == 0);
    const u32 total = (ntiles + 1) * p;
    u32 blocks = (total + 255) / 256;
    merge_bounds_kernel<<<blocks, 256, 0, stream>>>((const unsigned char *) recv, d, fast8, m,
                                                    (const u64 *) sorted_skeys, sorted_sid, ntiles, cut);
    CUDA_LAUNCH_CHECK();
    if (p == 2 && merge_rec16_ok(recv, out, elsize, offset, width, nwords)) {
        const int smem = MPSK_MERGE16_PADDED * 16 * 2;
        const u64 flip = is_signed ? (1ULL << 63) : 0ULL;
        cudaError_t e;
        if (offset == 8) {
            auto kern = merge_tile_rec16_kernel<true>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int) e;
            kern<<<ntiles, MPSK_MERGE16_THREADS, smem, stream>>>((const uint4 *) recv, flip, m, cut, (uint4 *) out, overflow);
        } else {
            auto kern = merge_tile_rec16_kernel<false>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int) e;
            kern<<<ntiles, MPSK_MERGE16_THREADS, smem, stream>>>((const uint4 *) recv, flip, m, cut, (uint4 *) out, overflow);
        }
        CUDA_LAUNCH_CHECK();
        return 0;
    }
    const uintptr_t a = ((uintptr_t) recv) | ((uintptr_t) out) | (uintptr_t) elsize;
    static int bucket = -1;
    if (bucket < 0) { const char * e = getenv("MPSORT_MERGE_BUCKET"); bucket = (e && atoi(e) > 0) ? 1 : 0; }
    if (bucket) {
        /* candidate: one-round bucket merge (16- and 8-byte pieces only; others keep the rounds) */
        const u64 * bk = (const u64 *) sorted_skeys;
        if ((a & 15) == 0) return launch_merge_bucket_tiles<uint4>(recv, d, fast8, m, cut, bk, out, overflow, ntiles, stream);
        if ((a & 7) == 0) return launch_merge_bucket_tiles<u64>(recv, d, fast8, m, cut, bk, out, overflow, ntiles, stream);
    }
    if ((a & 15) == 0) return launch_merge_tiles<uint4>(recv, d, fast8, m, cut, out, overflow, ntiles, stream);
    if ((a & 7) == 0) return launch_merge_tiles<u64>(recv, d, fast8, m, cut, out, overflow, ntiles, stream);
    if ((a & 3) == 0) return launch_merge_tiles<u32>(recv, d, fast8, m, cut, out, overflow, ntiles, stream);
    if ((a & 1) == 0) return launch_merge_tiles<unsigned short>(recv, d, fast8, m, cut, out, overflow, ntiles, stream);
    return launch_merge_tiles<unsigned char>(recv, d, fast8, m, cut, out, overflow, ntiles, stream);
}

/* ========================================================================= */
/* K6: record exchange by peer stores over NVLink                             */
/* ========================================================================= */
/*
 * Replaces MPI_Alltoallv / the sparse Isend-Irecv variant (mp-mpiu.c:69-236) inside
 * one box: every rank's receive buffer is mapped into all peers (CUDA IPC), and ONE
 * kernel per rank copies each destination's contiguous slice of the sorted records
 * straight into that peer's buffer with 16-byte stores (st.global on peer addresses
 * go over NVLink 5 / NVSwitch). Zero-length pairs cost nothing (the sparse variant).
 * CTAs are dealt to segments in proportion to their bytes.
 */
#define MPSK_P2P_MAX_SEGS 64
struct P2PPlan {
    const unsigned char * src[MPSK_P2P_MAX_SEGS];
    unsigned char * dst[MPSK_P2P_MAX_SEGS];
    unsigned long long bytes[MPSK_P2P_MAX_SEGS];
    u32 cta_begin[MPSK_P2P_MAX_SEGS + 1];     /* CTAs [cta_begin[k], cta_begin[k+1]) serve segment k */
    int nseg;
};

template <typename V>
__global__ void __launch_bounds__(512)
p2p_copy_kernel(P2PPlan plan)
{
    int k = 0;
    while (k + 1 < plan.nseg && blockIdx.x >= plan.cta_begin[k + 1]) k++;
    const u32 ncta = plan.cta_begin[k + 1] - plan.cta_begin[k];
    const u32 cta = blockIdx.x - plan.cta_begin[k];
    const V * __restrict__ src = (const V *) plan.src[k];
    V * __restrict__ dst = (V *) plan.dst[k];
    const size_t nv = plan.bytes[k] / sizeof(V);
    constexpr int U = 4;
    const size_t stride = (size_t) ncta * blockDim.x * U;
    for (size_t i0 = ((size_t) cta * blockDim.x) * U + threadIdx.x; i0 < nv; i0 += stride) {
        V v[U];
#pragma unroll
        for (int u = 0; u < U; u++) { const size_t i = i0 + (size_t) u * blockDim.x; if (i < nv) v[u] = src[i]; }
#pragma unroll
        for (int u = 0; u < U; u++) { const size_t i = i0 + (size_t) u * blockDim.x; if (i < nv) dst[i] = v[u]; }
    }
    __threadfence_system();        /* peer stores performed before the kernel is seen as done */
}

/* the same copy with the whole grid on ONE segment at a time, segments in the caller's
 * order (rotated by rank: the classic shifted all-to-all schedule, every GPU sends to
 * one peer and receives from one peer at a time) */
template <typename V>
__global__ void __launch_bounds__(512)
p2p_copy_seq_kernel(P2PPlan plan)
{
    constexpr int U = 4;
    const size_t stride = (size_t) gridDim.x * blockDim.x * U;
    for (int k = 0; k < plan.nseg; k++) {
        const V * __restrict__ src = (const V *) plan.src[k];
        V * __restrict__ dst = (V *) plan.dst[k];
        const size_t nv = plan.bytes[k] / sizeof(V);
        for (size_t i0 = ((size_t) blockIdx.x * blockDim.x) * U + threadIdx.x; i0 < nv; i0 += stride) {
            V v[U];
#pragma unroll
            for (int u = 0; u < U; u++) { const size_t i = i0 + (size_t) u * blockDim.x; if (i < nv) v[u] = src[i]; }
#pragma unroll
            for (int u = 0; u < U; u++) { const size_t i = i0 + (size_t) u * blockDim.x; if (i < nv) dst[i] = v[u]; }
        }
    }
    __threadfence_system();
}

extern "C" int mpsk_p2p_alltoallv(const void * const * src, void * const * dst, const uint64_t * bytes,
        const unsigned char * remote, int nseg, mpsk_stream_t stream)
{
    /* CTAs are dealt by bytes; weighting remote bytes higher did not help (profiles/r01_p2p_exchange.log) */
    static int wremote = -1, cta_mult = -1, sequential = -1;
    if (sequential < 0) sequential = getenv("MPSORT_P2P_SEQUENTIAL") ? 1 : 0;
    if (wremote < 0) { const char * e = getenv("MPSORT_P2P_REMOTE_WEIGHT"); wremote = e ? atoi(e) : 1; }
    if (cta_mult < 0) { const char * e = getenv("MPSORT_P2P_CTAS_PER_SM"); cta_mult = e ? atoi(e) : (nseg > 2 ? 1 : 4); }
    double weight[MPSK_P2P_MAX_SEGS];
    if (nseg > MPSK_P2P_MAX_SEGS) return (int) cudaErrorInvalidValue;
    P2PPlan plan;
    unsigned long long total = 0;
    uintptr_t align = 0;
    int n = 0;
    for (int k = 0; k < nseg; k++) {
        if (bytes[k] == 0) continue;
        plan.src[n] = (const unsigned char *) src[k];
        plan.dst[n] = (unsigned char *) dst[k];
        plan.bytes[n] = bytes[k];
        weight[n] = (double) bytes[k] * (remote[k] ? wremote : 1);
        total += (unsigned long long) weight[n];
        align |= (uintptr_t) src[k] | (uintptr_t) dst[k] | (uintptr_t) bytes[k];
        n++;
    }
    if (n == 0) return 0;
    plan.nseg = n;
    const u32 G = (u32) num_sms() * (u32) cta_mult;
    u32 acc = 0;
    for (int k = 0; k < n; k++) {
        u32 share = (u32) ((double) G * weight[k] / (double) total);
        if (share < 1) share = 1;
        plan.cta_begin[k] = acc;
        acc += share;
    }
    plan.cta_begin[n] = acc;
    cudaStream_t st = (cudaStream_t) stream;
    if (sequential) {
        if ((align & 15) == 0) p2p_copy_seq_kernel<uint4><<<G, 512, 0, st>>>(plan);
        else if ((align & 7) == 0) p2p_copy_seq_kernel<u64><<<G, 512, 0, st>>>(plan);
        else if ((align & 3) == 0) p2p_copy_seq_kernel<u32><<<G, 512, 0, st>>>(plan);
        else p2p_copy_seq_kernel<unsigned char><<<G, 512, 0, st>>>(plan);
        CUDA_LAUNCH_CHECK();
        return 0;
    }
    if ((align & 15) == 0) p2p_copy_kernel<uint4><<<acc, 512, 0, st>>>(plan);
    else if ((align & 7) == 0) p2p_copy_kernel<u64><<<acc, 512, 0, st>>>(plan);
    else if ((align & 3) == 0) p2p_copy_kernel<u32><<<acc, 512, 0, st>>>(plan);
    else p2p_copy_kernel<unsigned char><<<acc, 512, 0, st>>>(plan);
    CUDA_LAUNCH_CHECK();
    return 0;
}

/*
 * CANDIDATE, off by default (MPSORT_FUSED_PACK=1) and not yet run on a GPU: pack and exchange of
 * index mode in ONE kernel. gather_records_kernel writes the records in sorted order into a send
 * buffer and the slices then travel by DMA; here record i of the sorted order is read from
 * base[idx[i]] and stored straight into its destination rank's receive buffer (peer memory mapped
 * with CUDA IPC, st.global over NVLink; the own slice into the local receive buffer): one pass over
 * the records instead of two, no send buffer. Segments as in p2p_copy_kernel.
 */
struct P2PGatherPlan {
    const u32 * idx[MPSK_P2P_MAX_SEGS];          /* sorted-order source positions of the segment's records */
    unsigned char * dst[MPSK_P2P_MAX_SEGS];      /* where the segment's first record lands */
    unsigned long long nrec[MPSK_P2P_MAX_SEGS];
    u32 cta_begin[MPSK_P2P_MAX_SEGS + 1];
    int nseg;
};

template <typename V>
__global__ void __launch_bounds__(512)
p2p_gather_kernel(const V * __restrict__ base, P2PGatherPlan plan, u32 lpr /* V pieces per record */)
{
    int k = 0;
    while (k + 1 < plan.nseg && blockIdx.x >= plan.cta_begin[k + 1]) k++;
    const u32 ncta = plan.cta_begin[k + 1] - plan.cta_begin[k];
    const u32 cta = blockIdx.x - plan.cta_begin[k];
    const u32 * __restrict__ idx = plan.idx[k];
    V * __restrict__ dst = (V *) plan.dst[k];
    const size_t nv = (size_t) plan.nrec[k] * lpr;
    constexpr int U = 4;
    const size_t stride = (size_t) ncta * blockDim.x * U;
    for (size_t i0 = ((size_t) cta * blockDim.x) * U + threadIdx.x; i0 < nv; i0 += stride) {
        size_t src[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t i = i0 + (size_t) u * blockDim.x;
            src[u] = 0;
            if (i < nv) {
                const size_t rec = i / lpr;
                src[u] = (size_t) idx[rec] * lpr + (i - rec * lpr);
            }
        }
        V v[U];
#pragma unroll
        for (int u = 0; u < U; u++) { const size_t i = i0 + (size_t) u * blockDim.x; if (i < nv) v[u] = base[src[u]]; }
#pragma unroll
        for (int u = 0; u < U; u++) { const size_t i = i0 + (size_t) u * blockDim.x; if (i < nv) dst[i] = v[u]; }
    }
    __threadfence_system();        /* peer stores performed before the kernel is seen as done */
}

template <typename V>
static int launch_p2p_gather(const void * base, const P2PGatherPlan & plan, u32 grid, size_t elsize, cudaStream_t st)
{
    p2p_gather_kernel<V><<<grid, 512, 0, st>>>((const V *) base, plan, (u32) (elsize / sizeof(V)));
    CUDA_LAUNCH_CHECK();
    return 0;
}

extern "C" int mpsk_p2p_gather_alltoallv(const void * base, const uint32_t * const * idx, void * const * dst,
        const uint64_t * nrec, size_t elsize, int nseg, mpsk_stream_t stream)
{
    if (nseg > MPSK_P2P_MAX_SEGS || elsize == 0) return (int) cudaErrorInvalidValue;
    static int cta_mult = -1;
    if (cta_mult < 0) { const char * e = getenv("MPSORT_P2P_CTAS_PER_SM"); cta_mult = e ? atoi(e) : 2; if (cta_mult < 1) cta_mult = 1; }
    P2PGatherPlan plan;
    unsigned long long total = 0;
    uintptr_t align = (uintptr_t) base | (uintptr_t) elsize;
    int n = 0;
    for (int k = 0; k < nseg; k++) {
        if (nrec[k] == 0) continue;
        plan.idx[n] = idx[k];
        plan.dst[n] = (unsigned char *) dst[k];
        plan.nrec[n] = nrec[k];
        total += nrec[k];
        align |= (uintptr_t) dst[k];
        n++;
    }
    if (n == 0) return 0;
    plan.nseg = n;
    const u32 G = (u32) num_sms() * (u32) cta_mult;
    u32 acc = 0;
    for (int k = 0; k < n; k++) {
        u32 share = (u32) ((double) G * (double) plan.nrec[k] / (double) total);
        if (share < 1) share = 1;
        plan.cta_begin[k] = acc;
        acc += share;
    }
    plan.cta_begin[n] = acc;
    cudaStream_t st = (cudaStream_t) stream;
    if ((align & 15) == 0) return launch_p2p_gather<uint4>(base, plan, acc, elsize, st);
    if ((align & 7) == 0) return launch_p2p_gather<u64>(base, plan, acc, elsize, st);
    if ((align & 3) == 0) return launch_p2p_gather<u32>(base, plan, acc, elsize, st);
    if ((align & 1) == 0) return launch_p2p_gather<unsigned short>(base, plan, acc, elsize, st);
    return launch_p2p_gather<unsigned char>(base, plan, acc, elsize, st);
}

/* ========================================================================= */
/* K8: checksum                                                              */
/* ========================================================================= */

__global__ void __launch_bounds__(256)
checksum_kernel(const unsigned char * __restrict__ base, size_t nbytes, u64 * sum)
{
    /* head bytes up to 16-byte alignment, body as uint4 with dp4a, tail bytes */
    const uintptr_t addr = (uintptr_t) base;
    size_t head = (16 - (addr & 15)) & 15;
    if (head > nbytes) head = nbytes;
    const size_t nvec = (nbytes - head) / 16;
    const size_t tail_start = head + nvec * 16;
    const uint4 * body = (const uint4 *) (base + head);

    long long acc = 0;
    const size_t gtid = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t) gridDim.x * blockDim.x;
    int part = 0;
    int since = 0;
    for (size_t i = gtid; i < nvec; i += nthreads) {
        const uint4 v = body[i];
        part = __dp4a((int) v.x, 0x01010101, part);
        part = __dp4a((int) v.y, 0x01010101, part);
        part = __dp4a((int) v.z, 0x01010101, part);
        part = __dp4a((int) v.w, 0x01010101, part);
        if (++since == 65536) { acc += part; part = 0; since = 0; }
    }
    acc += part;
    if (gtid < head) acc += (signed char) base[gtid];
    if (gtid < nbytes - tail_start) acc += (signed char) base[tail_start + gtid];

#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, o);
    if ((threadIdx.x & 31) == 0 && acc != 0) atomicAdd(sum, (u64) acc);
}

extern "C" int mpsk_checksum(const void * base, size_t nbytes, uint64_t * sum, mpsk_stream_t stream)
{
    if (nbytes == 0) return 0;
    size_t blocks = (nbytes / 16 + 255) / 256;
    const size_t maxb = (size_t) num_sms() * 16;
    if (blocks > maxb) blocks = maxb;
    if (blocks == 0) blocks = 1;
    checksum_kernel<<<(unsigned) blocks, 256, 0, (cudaStream_t) stream>>>(
        (const unsigned char *) base, nbytes, (u64 *) sum);
    CUDA_LAUNCH_CHECK();
    return 0;
}

/* ========================================================================= */
/* bench / test support                                                      */
/* ========================================================================= */

__host__ __device__ __forceinline__ u64 mix64(u64 x)
{
    u64 z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* one synthetic record; the same arithmetic is restated in oracle/synth.h */
__device__ void synth_record(unsigned char * rec, size_t elsize, int kind, u64 seed,
                             u64 rank, u64 nranks, u64 n, u64 i)
{
    const u64 h = mix64(seed ^ (rank << 32) ^ i);
    const u64 tag = (rank << 40) + i;
    u64 key;
    if (kind == 1) {
        const u64 gi = rank * n + i;
        u64 src = gi;
        if (mix64(gi ^ 0xA5A5A5A5ULL) % 100 == 0 && n > 0) {
            src = (gi + 1 + mix64(gi ^ 0x5A5A5A5AULL) % n) % (nranks * n);
        }
        key = (src << 20) + (mix64(seed ^ src) & 0xFFFFFULL);
    } else if (kind == 2) {
        const double u = (double) (h >> 11) * (1.0 / 9007199254740992.0);
        const double u2 = u * u;
        const double u4 = u2 * u2;
        long long id = (long long) (u4 * 16777216.0) - (1LL << 20);
        if (mix64(h) % 20 == 0) id = 0;
        key = (u64) id;
    } else {
        key = h;
    }
    for (size_t b = 0; b < elsize; b++) {
        unsigned char v;
        if (b < 8) v = (unsigned char) (key >> (8 * b));
        else if (b < 16) v = (unsigned char) (tag >> (8 * (b - 8)));
        else v = (unsigned char) (mix64(h + b / 8) >> (8 * (b & 7)));
        rec[b] = v;
    }
}

__global__ void generate_kernel(unsigned char * dst, size_t n, size_t elsize, int kind, u64 seed,
                                u64 rank, u64 nranks)
{
    const size_t nthreads = (size_t) gridDim.x * blockDim.x;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        if (elsize == 16 && ((((uintptr_t) dst) & 15) == 0)) {
            __align__(16) unsigned char tmp[16];
            synth_record(tmp, 16, kind, seed, rank, nranks, n, i);
            ((uint4 *) dst)[i] = *(uint4 *) tmp;
        } else {
            synth_record(dst + i * elsize, elsize, kind, seed, rank, nranks, n, i);
        }
    }
}

extern "C" int mpsk_generate(void * dst, size_t n, size_t elsize, int kind, uint64_t seed,
        uint64_t rank, uint64_t nranks, mpsk_stream_t stream)
{
    if (n == 0) return 0;
    size_t blocks = (n + 255) / 256;
    const size_t maxb = (size_t) num_sms() * 16;
    if (blocks > maxb) blocks = maxb;
    generate_kernel<<<(unsigned) blocks, 256, 0, (cudaStream_t) stream>>>(
        (unsigned char *) dst, n, elsize, kind, seed, rank, nranks);
    CUDA_LAUNCH_CHECK();
    return 0;
}

__global__ void check_sorted_kernel(const unsigned char * __restrict__ base, size_t n, KeyDesc d, u32 nw,
                                    int check_ties, size_t tie_offset, u64 * violations, u64 * firstlast)
{
    const size_t nthreads = (size_t) gridDim.x * blockDim.x;
    u64 bad = 0;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
        const unsigned char * cur = base + i * d.elsize;
        if (i == 0 || i == n - 1) {
            for (u32 w = 0; w < nw; w++) {
                KeyDesc dd = d; dd.g = w;
                const u64 k = pack_key_word(cur, dd);
                if (i == 0) firstlast[w] = k;
                if (i == n - 1) firstlast[nw + w] = k;
            }
        }
        if (i == 0) continue;
        const unsigned char * prev = cur - d.elsize;
        int c = 0;
        for (int w = (int) nw - 1; w >= 0 && c == 0; w--) {
            KeyDesc dd = d; dd.g = (u32) w;
            const u64 a = pack_key_word(prev, dd);
            const u64 b = pack_key_word(cur, dd);
            c = (a > b) - (a < b);
        }
        if (c > 0) bad++;
        else if (c == 0 && check_ties) {
            const u64 ta = load_narrow(prev + tie_offset, 8);
            const u64 tb = load_narrow(cur + tie_offset, 8);
            if (ta >= tb) bad++;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(FULL_MASK, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(violations, bad);
}

extern "C" int mpsk_check_sorted(const void * base, size_t n, size_t elsize,
        size_t offset, uint32_t width, uint32_t nwords, int is_signed,
        int check_ties, size_t tie_offset,
        uint64_t * violations, uint64_t * firstlast, mpsk_stream_t stream)
{
    if (n == 0) return 0;
    KeyDesc d;
    d.elsize = elsize; d.offset = offset; d.width = width; d.nwords = nwords;
    d.is_signed = is_signed; d.g = 0; d.sub = 0;
    const u32 nw = (u32) (((size_t) width * nwords + 7) / 8);
    size_t blocks = (n + 255) / 256;
    const size_t maxb = (size_t) num_sms() * 16;
    if (blocks > maxb) blocks = maxb;
    check_sorted_kernel<<<(unsigned) blocks, 256, 0, (cudaStream_t) stream>>>(
        (const unsigned char *) base, n, d, nw, check_ties, tie_offset, (u64 *) violations, (u64 *) firstlast);
    CUDA_LAUNCH_CHECK();
    return 0;
}
