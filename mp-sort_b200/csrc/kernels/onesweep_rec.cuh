/* kernels/onesweep_rec.cuh -- K2 (record mode): onesweep pass over whole 8/16-byte records, TMA bulk stores; persistent candidate.
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
