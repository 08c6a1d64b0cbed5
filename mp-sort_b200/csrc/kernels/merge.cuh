/* kernels/merge.cuh -- K7: stable p-way merge of the received runs (samples, their merged order, bounds, tile kernels).
 * Part of the single translation unit mpsort_kernels.cu (included there, in order). */
/* ========================================================================= */
/* K7: stable p-way merge of the received runs (replaces the second radix_sort, */
/* mpsort-mpi.c:597, whose input is p sorted runs in source-rank order)          */
/* ========================================================================= */
/*
 * The receive buffer holds p sorted runs (run r = records [rdispl[r], rdispl[r+1])).
 * 1. merge_sample_kernel: every S-th key of every run (the last key of each full
 *    block of S) -> samples in (run, position) order.
 * 2. merge_rank_samples_kernel puts the samples in (key, run, position) order: every sample
 *    finds its rank by binary searches in the other runs' (sorted) sample lists.
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
#ifndef MPSK_MERGE_TILE
#define MPSK_MERGE_TILE 4096
#endif
#ifndef MPSK_MERGE_THREADS
#define MPSK_MERGE_THREADS 512
#endif
#ifndef MPSK_MERGE_MINBLOCKS
#define MPSK_MERGE_MINBLOCKS 2
#endif

struct MergeRuns {
    u32 p;
    u32 S;            /* sample stride */
    u32 k;            /* samples per tile */
    u32 rdispl[MPSK_MERGE_MAX_RUNS + 1];   /* run starts in records */
    u32 sstart[MPSK_MERGE_MAX_RUNS + 1];   /* first sample id of every run */
    /* ONE run may live elsewhere: the rank's own slice, read where the local sort left it (in the send
     * buffer) instead of being copied beside the received ones -- mostly sorted input keeps 99 % of its
     * records, and that copy was 1.3 ms of its 14.7 ms. Record i of run self_run is at
     * self_recv + (rdispl[self_run] + i) * elsize: the same indexing from another base. In the tile kernels
     * bit 31 of a source index says "from self_recv" (the host only asks for this below 2^31 records). */
    u32 self_run;                          /* MPSK_MERGE_NO_SELF: none */
    const unsigned char * self_recv;
};
#define MPSK_MERGE_NO_SELF 0xffffffffu
#define MPSK_MERGE_SELF_BIT 0x80000000u

__device__ __forceinline__ const unsigned char * merge_run_base(const unsigned char * recv, const MergeRuns & m, u32 r)
{
    return r == m.self_run ? m.self_recv : recv;
}

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
        skeys[s] = load_key_any(merge_run_base(recv, m, r) + pos * d.elsize, d, fast8);
    }
}

/*
 * 2. the merged order of the samples WITHOUT a sort: every run's samples are already sorted, so the
 * position of sample (run r, index j, key k) among all samples, ties by (run, index), is
 *      j + sum over r' < r of #{samples of r' with key <= k} + sum over r' > r of #{... with key < k},
 * p - 1 binary searches in arrays that sit in L2 (ns = n / S keys). One launch instead of the
 * eight-pass radix sort of the samples and its host round trip -- the fixed cost of every exchange part.
 */
__global__ void __launch_bounds__(256)
merge_rank_samples_kernel(const u64 * __restrict__ skeys, MergeRuns m, u64 * __restrict__ sorted_skeys,
                          u32 * __restrict__ sorted_sid)
{
    const u32 ns = m.sstart[m.p];
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < ns; s += gridDim.x * blockDim.x) {
        u32 r = 0;
        while (s >= m.sstart[r + 1]) r++;
        const u64 k = skeys[s];
        u32 rank = s - m.sstart[r];
        for (u32 q = 0; q < m.p; q++) {
            if (q == r) continue;
            const u64 * a = skeys + m.sstart[q];
            u32 lo = 0, hi = m.sstart[q + 1] - m.sstart[q];
            const bool upper = q < r;               /* lower runs win ties: count their equal keys too */
            while (lo < hi) {
                const u32 mid = lo + ((hi - lo) >> 1);
                const u64 v = a[mid];
                if (upper ? (v <= k) : (v < k)) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        sorted_skeys[rank] = k;
        sorted_sid[rank] = s;
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
                const unsigned char * base = merge_run_base(recv, m, r) + (size_t) m.rdispl[r] * d.elsize;
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

/* lane r < p asks the bulk-copy engine to bring sub-range r of tile tt into L2 (a hint: the range is
 * shrunk to 16-byte alignment on both sides) */
__device__ __forceinline__ void merge_prefetch_tile(const unsigned char * recv, size_t elsize, const MergeRuns & m,
        const u32 * __restrict__ cut, u32 tt, u32 ntiles, u32 r, u32 pf_dist)
{
    if (pf_dist == 0 || tt >= ntiles || r >= m.p) return;
    const u32 c0 = cut[tt * m.p + r], c1 = cut[(tt + 1) * m.p + r];
    if (c1 <= c0) return;
    const unsigned char * base = merge_run_base(recv, m, r);
    uintptr_t a0 = (uintptr_t) (base + ((size_t) m.rdispl[r] + c0) * elsize);
    uintptr_t a1 = (uintptr_t) (base + ((size_t) m.rdispl[r] + c1) * elsize);
    a0 = (a0 + 15) & ~(uintptr_t) 15;
    a1 &= ~(uintptr_t) 15;
    if (a1 <= a0) return;
    const u32 bytes = (u32) (a1 - a0);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a0), "r"(bytes) : "memory");
}

/* shared-memory index with one pad slot per 8 items: a thread's 8 consecutive outputs
 * are 64 bytes apart from its neighbour's, which would be a 16-way bank conflict */
#define MPD(i) ((i) + ((i) >> 3))
#define MPSK_MERGE_PADDED (MPSK_MERGE_TILE + MPSK_MERGE_TILE / 8)

/* FAST8: one aligned 8-byte key word (no generic key packing code in the kernel);
 * LPR1: a record is exactly one V (no division in the output loop) */
template <typename V, bool FAST8, bool LPR1>
__global__ void __launch_bounds__(MPSK_MERGE_THREADS, MPSK_MERGE_MINBLOCKS)
merge_tile_kernel(const unsigned char * __restrict__ recv, KeyDesc d, MergeRuns m,
                  const u32 * __restrict__ cut, unsigned char * __restrict__ out, u32 * __restrict__ overflow,
                  u32 ntiles, u32 pf_dist)
{
    constexpr int VT = MPSK_MERGE_TILE / MPSK_MERGE_THREADS;      /* items per thread */
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 * kA = (u64 *) smem_raw;
    u64 * kB = kA + MPSK_MERGE_PADDED;
    u32 * sA = (u32 *) (kB + MPSK_MERGE_PADDED);
    u32 * sB = sA + MPSK_MERGE_PADDED;
    __shared__ u32 seqoff[MPSK_MERGE_MAX_RUNS + 1];
    __shared__ u32 s_outstart;

    const u32 t = blockIdx.x, p = m.p, tid = threadIdx.x, lane = tid & 31u;
    /* EVERY warp fetches the 2p cut words (one lane per run, p <= 32; the same few words for all warps: L1
     * hits) and scans them for itself: nobody waits at a barrier for warp 0's two dependent global loads before
     * its own key loads can go out -- a fifth of all warp samples of the first version sat at that barrier
     * (ncu, profiles/r01_ncu_full_merge_tile_p8.csv). Warp 0 also leaves the offsets in shared memory for the
     * rounds; the barrier after the key staging publishes them. */
    u32 my_seqoff, my_srcbase, cnt, pe;
    {
        u32 c0 = 0, c1 = 0;
        if (lane < p) { c0 = cut[t * p + lane]; c1 = cut[(t + 1) * p + lane]; }
        const u32 len = c1 - c0;
        u32 incl = len, sum0 = c0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 y = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= (u32) o) incl += y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum0 += __shfl_xor_sync(FULL_MASK, sum0, o);
        my_seqoff = incl - len;                       /* lanes >= p hold the total: the run after the last starts there */
        my_srcbase = lane < p ? ((m.rdispl[lane] + c0) | (lane == m.self_run ? MPSK_MERGE_SELF_BIT : 0u)) : 0u;
        cnt = __shfl_sync(FULL_MASK, incl, 31);
        /* the rounds below merge only the runs that HAVE records in this tile (in run order: still stable):
         * ceil(log2) of their number instead of ceil(log2 p) rounds. Mostly sorted input sends a rank records
         * from itself and one neighbour -- one round where p = 8 would take three. */
        const u32 nonempty = __ballot_sync(FULL_MASK, lane < p && len > 0);
        pe = max((u32) __popc(nonempty), 1u);
        if (tid < 32) {
            if (lane < p && len > 0) seqoff[__popc(nonempty & ((1u << lane) - 1u))] = my_seqoff;
            if (tid == 0) {
                if (nonempty == 0) seqoff[0] = 0;
                seqoff[pe] = cnt;
                s_outstart = sum0;
            }
        } else if (tid < 64) {
            /* the p sub-ranges of tile t + pf_dist are asked into L2 (cp.async.bulk.prefetch.L2, one lane per
             * run): this kernel waits for its scattered key loads more than for anything else */
            merge_prefetch_tile(recv, d.elsize, m, cut, t + pf_dist, ntiles, tid - 32, pf_dist);
        }
    }
    if (cnt > MPSK_MERGE_TILE) {            /* cannot happen (tile bound); never corrupt memory */
        if (tid == 0) atomicAdd(overflow, 1u);
        return;
    }
    /* ---- load the keys of the p sub-ranges, run-major; all loads of a thread in flight */
    {
        u32 src[VT];
        u64 key[VT];
        u32 run[VT];
        /* the run of position i = the number of run starts at or below it (empty runs included), found with
         * warp-uniform broadcasts of the starts: every lane executes every shuffle */
#pragma unroll
        for (int k = 0; k < VT; k++) run[k] = 0;
        for (u32 q = 1; q < p; q++) {
            const u32 start = __shfl_sync(FULL_MASK, my_seqoff, q);
#pragma unroll
            for (int k = 0; k < VT; k++) run[k] += (tid + k * MPSK_MERGE_THREADS >= start) ? 1u : 0u;
        }
#pragma unroll
        for (int k = 0; k < VT; k++) {
            const u32 i = tid + k * MPSK_MERGE_THREADS;
            src[k] = __shfl_sync(FULL_MASK, my_srcbase, run[k]) + (i - __shfl_sync(FULL_MASK, my_seqoff, run[k]));
        }
#pragma unroll
        for (int k = 0; k < VT; k++) {
            const u32 i = tid + k * MPSK_MERGE_THREADS;
            if (i < cnt)
                key[k] = load_key_any(((src[k] & MPSK_MERGE_SELF_BIT) ? m.self_recv : recv)
                                      + (size_t) (src[k] & ~MPSK_MERGE_SELF_BIT) * d.elsize, d, FAST8);
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
    for (u32 w = 1; w < pe; w <<= 1) {
        u32 o = tid * VT;
        const u32 end = min(o + (u32) VT, cnt);
        u32 g = 0;                                  /* pair index: groups 2g and 2g+1 */
        while (o < end) {
            while (seqoff[min((2 * g + 2) * w, pe)] <= o) g++;
            const u32 a0 = seqoff[min(2 * g * w, pe)];
            const u32 a1 = seqoff[min((2 * g + 1) * w, pe)];
            const u32 b1 = seqoff[min((2 * g + 2) * w, pe)];
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
    const V * in_self = (const V *) m.self_recv;
    V * o = (V *) out + (size_t) s_outstart * lpr;
    const u32 totalv = cnt * lpr;
    for (u32 x0 = 0; x0 < totalv; x0 += VT * MPSK_MERGE_THREADS) {
        V v[VT];
#pragma unroll
        for (int k = 0; k < VT; k++) {
            const u32 x = x0 + tid + k * MPSK_MERGE_THREADS;
            if (x < totalv) {
                if (LPR1) {
                    const u32 s = sA[MPD(x)];
                    v[k] = ((s & MPSK_MERGE_SELF_BIT) ? in_self : in)[s & ~MPSK_MERGE_SELF_BIT];
                } else {
                    const u32 i = x / lpr, part = x - i * lpr;
                    const u32 s = sA[MPD(i)];
                    v[k] = ((s & MPSK_MERGE_SELF_BIT) ? in_self : in)[(size_t) (s & ~MPSK_MERGE_SELF_BIT) * lpr + part];
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
                        const u32 * __restrict__ cut, uint4 * __restrict__ out, u32 * __restrict__ overflow,
                        u32 ntiles, u32 pf_dist)
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
    } else if (tid < 64) {
        merge_prefetch_tile((const unsigned char *) recv, 16, m, cut, t + pf_dist, ntiles, tid - 32, pf_dist);
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
                rec[k] = (r == m.self_run ? (const uint4 *) m.self_recv : recv)[srcbase[r] + (i - seqoff[r])];
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
        const uint32_t * rdispl, const uint32_t * sstart, uint32_t self_run, const void * self_recv,
        uint64_t * skeys, mpsk_stream_t stream)
{
    if (p > MPSK_MERGE_MAX_RUNS) return (int) cudaErrorInvalidValue;
    MergeRuns m; m.p = p; m.S = S; m.k = k;
    m.self_run = self_recv ? self_run : MPSK_MERGE_NO_SELF; m.self_recv = (const unsigned char *) self_recv;
    for (u32 r = 0; r <= p; r++) { m.rdispl[r] = rdispl[r]; m.sstart[r] = sstart[r]; }
    const u32 ns = sstart[p];
    if (ns == 0) return 0;
    KeyDesc d; d.elsize = elsize; d.offset = offset; d.width = width; d.nwords = nwords; d.is_signed = is_signed; d.g = 0; d.sub = 0;
    const bool fast8 = (width == 8) && (nwords == 1) && (elsize % 8 == 0) && (offset % 8 == 0)
                       && (((((uintptr_t) recv) | ((uintptr_t) self_recv)) & 7) == 0);
    u32 blocks = (ns + 255) / 256;
    if (blocks > (u32) num_sms() * 8) blocks = (u32) num_sms() * 8;
    merge_sample_kernel<<<blocks, 256, 0, (cudaStream_t) stream>>>((const unsigned char *) recv, d, fast8, m, (u64 *) skeys);
    CUDA_LAUNCH_CHECK();
    return 0;
}

extern "C" int mpsk_merge_rank_samples(const uint64_t * skeys, uint32_t p, const uint32_t * sstart,
        uint64_t * sorted_skeys, uint32_t * sorted_sid, mpsk_stream_t stream)
{
    if (p > MPSK_MERGE_MAX_RUNS) return (int) cudaErrorInvalidValue;
    MergeRuns m; m.p = p; m.S = 0; m.k = 0; m.self_run = MPSK_MERGE_NO_SELF; m.self_recv = NULL;
    for (u32 r = 0; r <= p; r++) { m.rdispl[r] = 0; m.sstart[r] = sstart[r]; }
    const u32 ns = sstart[p];
    if (ns == 0) return 0;
    u32 blocks = (ns + 255) / 256;
    if (blocks > (u32) num_sms() * 8) blocks = (u32) num_sms() * 8;
    merge_rank_samples_kernel<<<blocks, 256, 0, (cudaStream_t) stream>>>((const u64 *) skeys, m, (u64 *) sorted_skeys, sorted_sid);
    CUDA_LAUNCH_CHECK();
    return 0;
}

/* MPSORT_PREFETCH_MERGE_TILES=d: tile + d is asked into L2 when a tile starts. Measured at 2^28 records
 * (profiles/r02_call3_predictor_prefetch_distances.log): 16-byte records 5.09 -> 4.89 ms at 8 runs, 3.76 -> 3.48 at 4,
 * 2.88 -> 2.78 at 2, with 74 .. 296; 48-byte records get SLOWER (4.29 -> 4.68 ms: eight 24 KB sub-ranges per tile
 * evict what the tiles at work still need), so the default is 148 up to 16-byte records and none above. */
#ifndef MPSK_MERGE_PREFETCH_TILES
#define MPSK_MERGE_PREFETCH_TILES 148
#endif
static u32 merge_prefetch_tiles(size_t elsize)
{
    static int v = -1;
    if (v < 0) { const char * e = getenv("MPSORT_PREFETCH_MERGE_TILES"); v = e ? atoi(e) : -2; }
    if (v >= 0) return (u32) v;
    return elsize <= 16 ? (u32) MPSK_MERGE_PREFETCH_TILES : 0u;
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
                                                           (unsigned char *) out, overflow, ntiles, merge_prefetch_tiles(d.elsize)); } while (0)
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
        const uint32_t * rdispl, const uint32_t * sstart, uint32_t self_run, const void * self_recv,
        const uint64_t * sorted_skeys, const uint32_t * sorted_sid, uint32_t ntiles,
        uint32_t * cut, uint32_t * overflow, mpsk_stream_t stream_)
{
    if (p > MPSK_MERGE_MAX_RUNS) return (int) cudaErrorInvalidValue;
    if (self_recv && (self_run >= p || rdispl[p] >= MPSK_MERGE_SELF_BIT)) return (int) cudaErrorInvalidValue;
    cudaStream_t stream = (cudaStream_t) stream_;
    MergeRuns m; m.p = p; m.S = S; m.k = k;
    m.self_run = self_recv ? self_run : MPSK_MERGE_NO_SELF; m.self_recv = (const unsigned char *) self_recv;
    for (u32 r = 0; r <= p; r++) { m.rdispl[r] = rdispl[r]; m.sstart[r] = sstart[r]; }
    KeyDesc d; d.elsize = elsize; d.offset = offset; d.width = width; d.nwords = nwords; d.is_signed = is_signed; d.g = 0; d.sub = 0;
    /* (every alignment question is asked of both bases) */
    const void * recv_probe = (const void *) (((uintptr_t) recv) | ((uintptr_t) self_recv));
    const bool fast8 = (width == 8) && (nwords == 1) && (elsize % 8 == 0) && (offset % 8 == 0) && ((((uintptr_t) recv_probe) & 7) == 0);
    const u32 total = (ntiles + 1) * p;
    u32 blocks = (total + 255) / 256;
    merge_bounds_kernel<<<blocks, 256, 0, stream>>>((const unsigned char *) recv, d, fast8, m,
                                                    (const u64 *) sorted_skeys, sorted_sid, ntiles, cut);
    CUDA_LAUNCH_CHECK();
    if (p == 2 && merge_rec16_ok(recv_probe, out, elsize, offset, width, nwords)) {
        const int smem = MPSK_MERGE16_PADDED * 16 * 2;
        const u64 flip = is_signed ? (1ULL << 63) : 0ULL;
        cudaError_t e;
        if (offset == 8) {
            auto kern = merge_tile_rec16_kernel<true>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int) e;
            kern<<<ntiles, MPSK_MERGE16_THREADS, smem, stream>>>((const uint4 *) recv, flip, m, cut, (uint4 *) out, overflow, ntiles, merge_prefetch_tiles(16));
        } else {
            auto kern = merge_tile_rec16_kernel<false>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int) e;
            kern<<<ntiles, MPSK_MERGE16_THREADS, smem, stream>>>((const uint4 *) recv, flip, m, cut, (uint4 *) out, overflow, ntiles, merge_prefetch_tiles(16));
        }
        CUDA_LAUNCH_CHECK();
        return 0;
    }
    const uintptr_t a = ((uintptr_t) recv_probe) | ((uintptr_t) out) | (uintptr_t) elsize;
    if ((a & 15) == 0) return launch_merge_tiles<uint4>(recv, d, fast8, m, cut, out, overflow, ntiles, stream);
    if ((a & 7) == 0) return launch_merge_tiles<u64>(recv, d, fast8, m, cut, out, overflow, ntiles, stream);
    if ((a & 3) == 0) return launch_merge_tiles<u32>(recv, d, fast8, m, cut, out, overflow, ntiles, stream);
    if ((a & 1) == 0) return launch_merge_tiles<unsigned short>(recv, d, fast8, m, cut, out, overflow, ntiles, stream);
    return launch_merge_tiles<unsigned char>(recv, d, fast8, m, cut, out, overflow, ntiles, stream);
}
