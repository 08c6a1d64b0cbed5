/* kernels/onesweep_rec.cuh -- K2 (record mode): onesweep pass over whole 8/16-byte records, TMA bulk stores.
 * Part of the single translation unit mpsort_kernels.cu (included there, in order). */
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
/* tile + 148 is asked into L2 when a tile starts (one CTA wave of 3 per SM is 444 tiles): a record pass
 * takes 1.91 ms without, 1.78 ms with 148 or 296, 1.80 with 444, 1.87 with 592 and 2.05 with 888 and more
 * (profiles/r02_call2_static_prefetch_fixup_merge.log) */
#ifndef MPSK_REC_PREFETCH_TILES
#define MPSK_REC_PREFETCH_TILES 148
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

/*
 * TICKET = false (default): tile = blockIdx.x. A tile only ever waits for tiles with smaller
 * numbers, and CTAs of a 1-D grid are dispatched in blockIdx order, so every tile a resident CTA
 * waits for is resident or done -- the assumption CUB's decoupled look-back scan makes too
 * (cub/agent/agent_scan.cuh: tile_idx = start_tile + blockIdx.x). It takes one global atomic
 * round trip and one barrier off the front of every tile: 1.99 -> 1.90 ms per pass
 * (profiles/r02_call1_tests_candidates_bench_n1.log). TICKET = true (MPSORT_TICKET_TILES=1) hands
 * tiles out by an atomic counter instead, which needs no assumption about the dispatch order.
 * pf_dist > 0: one thread asks the bulk-copy engine to bring tile + pf_dist into L2
 * (cp.async.bulk.prefetch.L2): by the time that tile's CTA starts, its loads hit L2.
 */
template <int THREADS, int IPT, typename ITEM, bool TICKET>
__global__ void __launch_bounds__(THREADS, MPSK_REC_MINBLOCKS)
onesweep_rec_kernel(const ITEM * __restrict__ in, ITEM * __restrict__ out,
                      u32 n, u32 shift, u32 khi, u64 flip, const u32 * __restrict__ bins,
                      LookbackBufs lb, u32 * ticket, u32 pf_dist)
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

    u32 tile = blockIdx.x;
    if (TICKET) {
        if (tid == 0) s_misc[0] = atomicAdd(ticket, 1u);
        __syncthreads();
        tile = s_misc[0];
    }
    if (pf_dist && tid == 0) {
        const u64 first = ((u64) tile + pf_dist) * (u64) TILE;
        if (first < (u64) n) {
            const u64 left = ((u64) n - first) * sizeof(ITEM);
            const u32 bytes = (u32) (left < (u64) TILE * sizeof(ITEM) ? (left & ~15ULL) : (u64) TILE * sizeof(ITEM));
            if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(in + first), "r"(bytes) : "memory");
        }
    }
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
    /* every warp zeroes its own histogram while its loads are in flight */
#pragma unroll
    for (int k = 0; k < 8; k++) my_hist[lane + 32 * k] = 0;
    __syncwarp();
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

static u32 sweep_prefetch_tiles()
{
    static int v = -1;
    if (v < 0) { const char * e = getenv("MPSORT_PREFETCH_TILES"); v = e ? atoi(e) : MPSK_REC_PREFETCH_TILES; if (v < 0) v = 0; }
    return (u32) v;
}

template <typename ITEM>
static int launch_rec_pass(const void * in, void * out, size_t n, int shift, int key_in_high, uint64_t flip,
                           const uint32_t * bins, void * scratch, cudaStream_t stream)
{
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
    if (sweep_ticket_tiles()) {
        auto kern = onesweep_rec_kernel<MPSK_REC_THREADS, IPT, ITEM, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return (int) e;
        kern<<<(unsigned) ntiles, MPSK_REC_THREADS, Cfg::SMEM, stream>>>(
            (const ITEM *) in, (ITEM *) out, (u32) n, (u32) shift, key_in_high ? 1u : 0u, (u64) flip,
            bins, lb, ticket, sweep_prefetch_tiles());
    } else {
        auto kern = onesweep_rec_kernel<MPSK_REC_THREADS, IPT, ITEM, false>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return (int) e;
        kern<<<(unsigned) ntiles, MPSK_REC_THREADS, Cfg::SMEM, stream>>>(
            (const ITEM *) in, (ITEM *) out, (u32) n, (u32) shift, key_in_high ? 1u : 0u, (u64) flip,
            bins, lb, ticket, sweep_prefetch_tiles());
    }
    CUDA_LAUNCH_CHECK();
    return 0;
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
