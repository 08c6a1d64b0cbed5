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
