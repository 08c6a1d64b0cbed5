/* kernels/splitter.cuh -- K4: splitter count / select / final; the peer-memory descent kernel; u64 sum.
 * Part of the single translation unit mpsort_kernels.cu (included there, in order). */
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
 * The whole byte-wise descent in ONE kernel per GPU, the per-level sums taken over peer memory instead of
 * count kernel + ncclAllReduce + select kernel per level (the default of one process per GPU).
 *
 * Every rank owns a mailbox in device memory that all peers have mapped (CUDA IPC; plain pointers for rank
 * threads of one process): words[2 parities][source rank][splitter][256 digits]. Block b works on splitter b on
 * every rank. Per level thread d counts the local keys <= candidate d (a binary search, from the second level on
 * inside the range the previous level left: the searches shrink 256-fold per level), PUSHES the count into the
 * mailbox of every peer as ONE 64-bit word {tag = seq + level + 1 : count} and then polls ITS OWN mailbox (local
 * memory) for the words of the peers with that tag. Data and flag are one word, so no fence and no separate flag
 * round trip is needed, and nothing is read over NVLink: a level costs one remote-store latency. All ranks compute
 * the same sums, so nothing is broadcast. Two parities suffice: a rank reaches level L+2 only after every peer
 * published level L+1, which a peer does after it has read level L.
 * The first version (pull: counts in the own mailbox, a release flag per splitter, two system fences per level,
 * peers' counts read over NVLink) took 0.33 ms for eight levels at 8 GPUs (profiles/r02_call_n8_final.log).
 * Block b only ever waits for block b of the peers' kernels; <= 63 blocks are always co-resident. A wait that
 * exceeds `timeout` clock cycles sets *err and leaves (the host aborts the job) instead of hanging the GPU.
 */
#define MPSK_PEER_MAXS 63
#define MPSK_PEER_MAXR 64
struct PeerBoxes { unsigned long long * box[64]; };
__host__ __device__ constexpr size_t peer_box_count_words() { return (size_t) 2 * MPSK_PEER_MAXR * MPSK_PEER_MAXS * 256; }

__device__ __forceinline__ u64 ld_relaxed_sys_u64(const u64 * p)
{
    u64 v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(u64 * p, u64 v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

/* number of keys <= cand, known to lie in [lo, hi] */
__device__ __forceinline__ u64 bound_key_in(const mpsk_keyview & v, size_t lo, size_t hi, const u64 * cand, u32 nw)
{
    while (lo < hi) {
        const size_t mid = lo + ((hi - lo) >> 1);
        if (cmp_key(v, mid, cand, nw) <= 0) lo = mid + 1; else hi = mid;
    }
    return (u64) lo;
}

__global__ void __launch_bounds__(256)
splitter_descent_peer_kernel(mpsk_keyview v, size_t n, u32 nw, u64 * __restrict__ prefix, const u64 * __restrict__ target,
                             int level0, int nlevels, u32 me, u32 p, PeerBoxes boxes, u32 seq,
                             long long timeout, u32 * __restrict__ err)
{
    __shared__ u64 s_prefix[MPSK_MAX_KEY_WORDS];
    __shared__ u64 s_cnt[257];                 /* [d + 1] = local keys <= candidate d; [0] = below the whole range */
    __shared__ u32 s_min, s_abort;
    const u32 b = blockIdx.x, d = threadIdx.x;
    if (d < nw) s_prefix[d] = prefix[(size_t) b * nw + d];
    if (d == 0) s_abort = 0;
    __syncthreads();
    const u64 tgt = target[b];
    u64 * mybox = boxes.box[me];
    size_t lo = 0, hi = n;                     /* where the local keys with the prefix decided so far sit */
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
        const u64 c = bound_key_in(v, lo, hi, cand, nw);
        s_cnt[d + 1] = c;
        if (d == 0) { s_cnt[0] = (u64) lo; s_min = 255u; }
        /* push {tag : count} to every peer, then gather the peers' words from my own mailbox */
        const u32 want = seq + (u32) level + 1u;
        const u64 word = ((u64) want << 32) | (u64) (u32) c;
        const size_t myslot = (((size_t) par * MPSK_PEER_MAXR + me) * MPSK_PEER_MAXS + b) * 256 + d;
        for (u32 r = 0; r < p; r++)
            if (r != me) st_relaxed_sys_u64(boxes.box[r] + myslot, word);
        u64 sum = c;
        const long long t0 = clock64();
        for (u32 r = 0; r < p; r++) {
            if (r == me) continue;
            const u64 * src = mybox + (((size_t) par * MPSK_PEER_MAXR + r) * MPSK_PEER_MAXS + b) * 256 + d;
            u64 w = ld_relaxed_sys_u64(src);
            while ((u32) (w >> 32) != want) {
                if (clock64() - t0 > timeout) { s_abort = 1; break; }
                __nanosleep(40);
                w = ld_relaxed_sys_u64(src);
            }
            sum += (u64) (u32) w;
        }
        __syncthreads();
        if (s_abort) {
            if (d == 0) atomicExch(err, 1u);
            return;
        }
        if (sum >= tgt) atomicMin(&s_min, d);
        __syncthreads();
        const u32 pick = s_min;
        lo = pick == 0 ? (size_t) s_cnt[0] : (size_t) s_cnt[pick];        /* local keys <= candidate pick - 1 */
        hi = (size_t) s_cnt[pick + 1];
        if (d == 0) s_prefix[wi] |= ((u64) pick) << sh;
        __syncthreads();
    }
    if (d < nw) prefix[(size_t) b * nw + d] = s_prefix[d];
}

extern "C" size_t mpsk_peer_box_bytes(void) { return peer_box_count_words() * sizeof(u64); }

extern "C" int mpsk_splitter_descent_peer(struct mpsk_keyview view, size_t n, uint32_t nw,
        uint64_t * prefix, const uint64_t * target, int nsplit, int level0, int nlevels,
        uint32_t me, uint32_t p, void * const * boxes, uint32_t seq, uint32_t * err, mpsk_stream_t stream)
{
    if (nsplit <= 0 || level0 >= nlevels) return 0;
    if (nw > MPSK_MAX_KEY_WORDS || nsplit > MPSK_PEER_MAXS || p > MPSK_PEER_MAXR || me >= p || n > 0xffffffffu) return (int) cudaErrorInvalidValue;
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
