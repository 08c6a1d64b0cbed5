/* kernels/hybrid_fixup.cuh -- hybrid sort: run fix-up after the high-digit passes, predictor sampling.
 * Part of the single translation unit mpsort_kernels.cu (included there, in order). */
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

#ifndef FIX_MINBLOCKS
#define FIX_MINBLOCKS 8
#endif
/* tile + 444 is asked into L2 when a tile starts: 0.92 -> 0.83 ms per 2^28 records with 296 or 592,
 * 0.87 with 1184, 1.28 with 2368 (profiles/r02_call3_predictor_prefetch_distances.log) */
#ifndef FIX_PREFETCH_TILES
#define FIX_PREFETCH_TILES 444
#endif

/* what one entry of the compacted list (a position that is not a run of its own) has to do:
 * returns the position its record moves to (0xffffffff: it stays) and reads the record */
template <typename ITEM>
__device__ __forceinline__ u32 fixup_entry(const ITEM * __restrict__ recs, size_t t0, u32 i, u32 cnt,
        const u64 * s_key, const u32 * s_head, int nwords, u64 lomask, u32 lobits,
        u32 * __restrict__ worklist, u32 * __restrict__ nwork, u32 cap, ITEM & moved)
{
    (void) lobits;
    /* run start: last head at or before i */
    int w = (int) (i >> 5);
    u32 bits = s_head[w] & (0xffffffffu >> (31 - (i & 31)));
    while (bits == 0 && w > 0) { w--; bits = s_head[w]; }
    if (bits == 0) return 0xffffffffu;                    /* continuation of a run owned by an earlier tile */
    const u32 rs = (u32) w * 32 + (31 - __clz(bits));
    if (rs >= (u32) FIX_T) return 0xffffffffu;            /* head lies in the look-ahead: the next tile owns it */
    /* run end: first head after i (the sentinel counts) */
    u32 w2 = (i + 1) >> 5;
    u32 b2 = s_head[w2] & (0xffffffffu << ((i + 1) & 31));
    while (b2 == 0 && w2 + 1 < (u32) nwords && (w2 + 1) * 32 <= cnt + 31) { w2++; b2 = s_head[w2]; }
    const u32 re = b2 ? (w2 * 32 + (__ffs(b2) - 1)) : 0xffffffffu;
    if (re == 0xffffffffu || re > cnt || re - rs > (u32) FIX_HALO) {
        /* too long for this kernel: the run head reports it */
        if (i == rs) {
            const u32 slot = atomicAdd(nwork, 1u);
            if (slot < cap) worklist[slot] = (u32) (t0 + rs);
        }
        return 0xffffffffu;
    }
    if (re - rs < 2) return 0xffffffffu;
    const u64 mine = s_key[i + 1] & lomask;
    u32 rank = 0;
    for (u32 j = rs; j < re; j++) {
        const u64 other = s_key[j + 1] & lomask;
        rank += (other < mine) || (other == mine && j < i);
    }
    if (rs + rank == i) return 0xffffffffu;
    moved = recs[t0 + i];
    return rs + rank;
}

/* bit b set: position 32 w + b of the tile is inside the data and NOT a run of its own */
__device__ __forceinline__ u32 fixup_votes(const u32 * s_head, u32 cnt, u32 w)
{
    const u32 H = s_head[w], Hn = s_head[w + 1];
    const u32 single = H & ((H >> 1) | (Hn << 31));
    const u32 first = w * 32;
    const u32 inside = first >= cnt ? 0u : (cnt - first >= 32 ? 0xffffffffu : ((1u << (cnt - first)) - 1u));
    return ~single & inside;
}

template <typename ITEM, bool KHI>
__global__ void __launch_bounds__(FIX_THREADS, FIX_MINBLOCKS)
fixup_rec_kernel(ITEM * __restrict__ recs, u32 n, u64 flip, u32 lobits,
                   u32 * __restrict__ worklist, u32 * __restrict__ nwork, u32 cap, u32 pf_dist, u32 tile0)
{
    constexpr int CAP = FIX_T + FIX_HALO;
    constexpr int NW = CAP / 32;                  /* head words that describe positions of this tile */
    constexpr int WORDS = NW + 1;
    constexpr int NLD = CAP / FIX_THREADS;
    static_assert(CAP % FIX_THREADS == 0, "tile + halo must be a multiple of the block size");
    static_assert(NW == NLD * (FIX_THREADS / 32), "one warp row per head word");
    static_assert(NW <= 96, "the list offsets are scanned by one warp, three words per lane");
    static_assert(FIX_HALO <= FIX_THREADS, "a run must not span more than two rounds of the move phase");
    static_assert(CAP < 65536, "list entries are 16-bit positions");
    /* Only the keys are staged: the few records that move are re-read from global. 23 KB of shared
     * memory and <= 32 registers: eight CTAs per SM keep enough key loads in flight for what is, for
     * random keys, ONE read of the array (6 % of the records move). The first version held a moved
     * record per load round in registers (64 of them, four CTAs per SM) and ran at 54 % of the HBM rate
     * of a plain read (1.22 ms per 2^28 records, profiles/r02_call1_tests_candidates_bench_n1.log). */
    __shared__ u64 s_key[CAP + 1];            /* [0] = key of the record before the tile */
    __shared__ u32 s_head[WORDS];
    __shared__ u32 s_off[96];                 /* per head word: positions on the list before it */
    __shared__ unsigned short s_list[CAP];
    __shared__ u32 s_nlist;

    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31u;
    const size_t t0 = ((size_t) blockIdx.x + tile0) * FIX_T;
    const u32 avail = (u32) ((size_t) n - t0);
    const u32 cnt = avail < (u32) CAP ? avail : (u32) CAP;
    const bool at_end = (t0 + cnt == n);
    const u64 lomask = lobits >= 64 ? ~0ULL : ((1ULL << lobits) - 1ULL);
    constexpr u32 W = sizeof(ITEM) / 8;                            /* u64 words per record */
    const u64 * keys = (const u64 *) recs + (KHI ? 1 : 0);       /* key of record i at keys[W*i] */

    /* tile + pf_dist is asked into L2 while this one is worked on (cp.async.bulk.prefetch.L2) */
    if (pf_dist && tid == 0) {
        const u64 first = ((u64) blockIdx.x + tile0 + pf_dist) * (u64) FIX_T;
        if (first < (u64) n) {
            const u64 left = ((u64) n - first) * sizeof(ITEM);
            const u32 bytes = (u32) (left < (u64) FIX_T * sizeof(ITEM) ? (left & ~15ULL) : (u64) FIX_T * sizeof(ITEM));
            if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(recs + first), "r"(bytes) : "memory");
        }
    }
    {
        u64 tmp[NLD];
#pragma unroll
        for (int k = 0; k < NLD; k++) {
            const u32 i = tid + k * FIX_THREADS;
            if (i < cnt) tmp[k] = keys[W * (t0 + i)];
        }
        if (tid == 0) s_key[0] = t0 ? (keys[W * (t0 - 1)] ^ flip) : 0ULL;
#pragma unroll
        for (int k = 0; k < NLD; k++) {
            const u32 i = tid + k * FIX_THREADS;
            if (i < cnt) s_key[i + 1] = tmp[k] ^ flip;
        }
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
            if (lane == 0) s_head[i >> 5] = word;
        }
    }
    __syncthreads();
    /* ---- the positions that are NOT a run of their own (6 % for random keys) are compacted into a list
     * IN POSITION ORDER, so that the expensive part below runs with full warps. Position i is a run of its
     * own when head bits i and i+1 are both set: a whole word of positions is decided by two loads and a
     * few word operations. 1: count per word; 2: one warp scans the counts; 3: write the list. */
    if (tid < 32) {
        u32 c[3], t = 0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const u32 w = 3 * tid + j;
            c[j] = w < (u32) NW ? (u32) __popc(fixup_votes(s_head, cnt, w)) : 0u;
            t += c[j];
        }
        u32 incl = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 y = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= (u32) o) incl += y;
        }
        u32 run = incl - t;
#pragma unroll
        for (int j = 0; j < 3; j++) { s_off[3 * tid + j] = run; run += c[j]; }
        if (tid == 31) s_nlist = incl;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NLD; k++) {
        const u32 w = (tid >> 5) + k * (FIX_THREADS / 32);
        const u32 votes = fixup_votes(s_head, cnt, w);
        if ((votes >> lane) & 1u) s_list[s_off[w] + __popc(votes & lanemask_lt())] = (unsigned short) (w * 32 + lane);
    }
    __syncthreads();
    /* ---- the move phase, in rounds of FIX_THREADS list entries (one round for random keys). A run's
     * positions are consecutive on the list and there are at most FIX_HALO <= FIX_THREADS of them, so a
     * run spans at most two consecutive rounds: the records of round r+1 are read before the barrier that
     * precedes the writes of round r, and whatever round r writes belongs to runs that rounds r-1, r and
     * r+1 have read by then. Two moved records per thread instead of one per load round. */
    const u32 nlist = s_nlist;
    ITEM mv_a, mv_b;
    u32 tgt_a = 0xffffffffu;
    if (tid < nlist)
        tgt_a = fixup_entry<ITEM>(recs, t0, s_list[tid], cnt, s_key, s_head, WORDS, lomask, lobits, worklist, nwork, cap, mv_a);
    for (u32 base = 0; base < nlist; base += FIX_THREADS) {
        u32 tgt_b = 0xffffffffu;
        const u32 e = base + FIX_THREADS + tid;
        if (e < nlist)
            tgt_b = fixup_entry<ITEM>(recs, t0, s_list[e], cnt, s_key, s_head, WORDS, lomask, lobits, worklist, nwork, cap, mv_b);
        __syncthreads();
        if (tgt_a != 0xffffffffu) recs[t0 + tgt_a] = mv_a;
        tgt_a = tgt_b;
        mv_a = mv_b;
    }
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

extern "C" size_t mpsk_fixup_tile_items(void) { return FIX_T; }

extern "C" int mpsk_fixup_rec(void * recs, size_t n, size_t elsize, int key_in_high, uint64_t flip, uint32_t lobits,
        uint32_t * worklist, uint32_t * nwork, uint32_t cap, size_t tile0, size_t ntiles, mpsk_stream_t stream)
{
    if (n == 0) return 0;
    const size_t all = (n + FIX_T - 1) / FIX_T;
    if (tile0 >= all) return 0;
    size_t tiles = all - tile0;
    if (ntiles != 0 && ntiles < tiles) tiles = ntiles;
    static int pf = -1;                  /* MPSORT_PREFETCH_FIXUP_TILES: eight CTAs per SM, a wave is 1184 tiles */
    if (pf < 0) { const char * e = getenv("MPSORT_PREFETCH_FIXUP_TILES"); pf = e ? atoi(e) : FIX_PREFETCH_TILES; if (pf < 0) pf = 0; }
    if (elsize == 8)
        fixup_rec_kernel<u64, false><<<(unsigned) tiles, FIX_THREADS, 0, (cudaStream_t) stream>>>(
            (u64 *) recs, (u32) n, (u64) flip, lobits, worklist, nwork, cap, (u32) pf, (u32) tile0);
    else if (key_in_high)
        fixup_rec_kernel<uint4, true><<<(unsigned) tiles, FIX_THREADS, 0, (cudaStream_t) stream>>>(
            (uint4 *) recs, (u32) n, (u64) flip, lobits, worklist, nwork, cap, (u32) pf, (u32) tile0);
    else
        fixup_rec_kernel<uint4, false><<<(unsigned) tiles, FIX_THREADS, 0, (cudaStream_t) stream>>>(
            (uint4 *) recs, (u32) n, (u64) flip, lobits, worklist, nwork, cap, (u32) pf, (u32) tile0);
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

/* predictor: equal PAIRS among the high parts of `s` evenly spaced records, for up to two values of lobits
 * at once. Every sample is inserted into an open-addressing hash table (keys + counts, in L2): the count its
 * slot held before is the number of equal samples inserted earlier, and those sum to k(k-1)/2 per value
 * whatever the order. One launch; the sort of the samples that used to do this took a dozen and a host
 * round trip. */
struct PredArgs { u32 lobits[2]; u32 nl; };

/* one insertion: the number of equal values inserted before this one */
__device__ __forceinline__ u64 pred_insert(unsigned long long * tk, unsigned long long * tc, u64 tsize, u64 value)
{
    const u64 mask = tsize - 1;
    const u64 v = value + 1ULL;                                   /* 0 = empty slot */
    u64 h = mset_hash64(v) & mask;
    for (u64 probe = 0; probe < tsize; probe++) {
        const unsigned long long prev = atomicCAS(&tk[h], 0ULL, (unsigned long long) v);
        if (prev == 0ULL || prev == v) return atomicAdd(&tc[h], 1ULL);
        h = (h + 1) & mask;
    }
    return 0;
}

__global__ void __launch_bounds__(256)
prefix_pairs_kernel(const u64 * __restrict__ recs, u32 W, size_t n, u32 s, u32 khi, u64 flip, PredArgs a,
                    u64 * __restrict__ table, u32 log2t, u64 * __restrict__ pairs)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    const u64 tsize = 1ULL << log2t;
    u64 found[3] = { 0, 0, 0 };
    if (i < s) {
        /* pseudo-random positions over the WHOLE array: the estimate counts samples that fall into the same run,
         * and evenly spaced (or one-per-stratum) samples of nearly sorted keys never do -- they would report
         * singleton runs for keys whose runs are 16 long. Positions drawn twice are counted too (table nl) and
         * taken off by the caller: pairs[nl] of them. */
        const size_t pos = (size_t) (mset_hash64(0x5EED5A3Bu + i) % (u64) n);
        const u64 key = rec_key_at(recs, pos, W, khi, flip);
        for (u32 j = 0; j <= a.nl; j++) {
            unsigned long long * tk = (unsigned long long *) table + (size_t) j * 2 * tsize;
            const u64 value = j == a.nl ? (u64) pos : (a.lobits[j] >= 64 ? 0ULL : key >> a.lobits[j]);
            found[j] = pred_insert(tk, tk + tsize, tsize, value);
        }
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
        u64 v = found[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd((unsigned long long *) &pairs[j], (unsigned long long) v);
    }
}

extern "C" int mpsk_prefix_pairs(const void * recs, size_t n, size_t elsize, uint32_t s, int key_in_high, uint64_t flip,
        const uint32_t * lobits, uint32_t nl, uint64_t * table, uint32_t log2_tsize, uint64_t * pairs, mpsk_stream_t stream)
{
    if (s == 0 || nl == 0 || n == 0) return 0;
    if (nl > 2 || ((size_t) 1 << log2_tsize) < 2 * (size_t) s) return (int) cudaErrorInvalidValue;
    PredArgs a;
    a.nl = nl; a.lobits[0] = lobits[0]; a.lobits[1] = nl > 1 ? lobits[1] : 0;
    prefix_pairs_kernel<<<(s + 255) / 256, 256, 0, (cudaStream_t) stream>>>(
        (const u64 *) recs, (u32) (elsize / 8), n, s, (key_in_high && elsize == 16) ? 1u : 0u, (u64) flip, a,
        (u64 *) table, log2_tsize, (u64 *) pairs);
    CUDA_LAUNCH_CHECK();
    return 0;
}
