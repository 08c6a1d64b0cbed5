/* kernels/onesweep.cuh -- K2: onesweep pass over (u64 key, u32 index) pairs: ranking, two-level decoupled look-back.
 * Part of the single translation unit mpsort_kernels.cu (included there, in order). */
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
template <int THREADS, int IPT, bool IOTA, bool TICKET>
__global__ void __launch_bounds__(THREADS, MPSK_SWEEP_MINBLOCKS)
onesweep_kernel(const u64 * __restrict__ kin, const u32 * __restrict__ vin,
                u64 * __restrict__ kout, u32 * __restrict__ vout,
                u32 n, u32 shift, const u32 * __restrict__ bins,
                LookbackBufs lb, u32 * ticket, u32 pf_dist)
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

    /* tile = blockIdx.x unless tickets were asked for (see onesweep_rec_kernel) */
    u32 tile = blockIdx.x;
    if (TICKET) {
        if (tid == 0) s_misc[0] = atomicAdd(ticket, 1u);
        __syncthreads();
        tile = s_misc[0];
    }
    /* tile + pf_dist is asked into L2 (see onesweep_rec_kernel): its keys, and its values unless they are 0..n-1 */
    if (pf_dist && tid == 0) {
        const u64 first = ((u64) tile + pf_dist) * (u64) TILE;
        if (first < (u64) n) {
            const u64 left = (u64) n - first, items = left < (u64) TILE ? left : (u64) TILE;
            const u32 kb = (u32) ((items * 8) & ~15ULL), vb = (u32) ((items * 4) & ~15ULL);
            if (kb) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(kin + first), "r"(kb) : "memory");
            if (!IOTA && vb) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(vin + first), "r"(vb) : "memory");
        }
    }
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
    /* every warp zeroes its own histogram while its loads are in flight */
#pragma unroll
    for (int k = 0; k < 8; k++) my_hist[lane + 32 * k] = 0;
    __syncwarp();
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

/* run-time switches of the onesweep passes (read once): MPSORT_TICKET_TILES=1 hands tiles out by an
 * atomic ticket instead of blockIdx.x; MPSORT_PREFETCH_TILES=d (record passes) prefetches tile + d into L2 */
static int sweep_ticket_tiles()
{
    static int v = -1;
    if (v < 0) { const char * e = getenv("MPSORT_TICKET_TILES"); v = (e && atoi(e) > 0) ? 1 : 0; }
    return v;
}

/* index passes: MPSORT_PREFETCH_INDEX_TILES (tiles are 6144 pairs, two CTAs per SM: a wave is 296 tiles) */
/* 1.59 ms per pass without, 1.50 with 74 or 148, 1.52 with 296, 1.65 with 592 (2^28 pairs,
 * profiles/r02_call3_predictor_prefetch_distances.log) */
#ifndef MPSK_SWEEP_PREFETCH_TILES
#define MPSK_SWEEP_PREFETCH_TILES 148
#endif
static u32 sweep_prefetch_index_tiles()
{
    static int v = -1;
    if (v < 0) { const char * e = getenv("MPSORT_PREFETCH_INDEX_TILES"); v = e ? atoi(e) : MPSK_SWEEP_PREFETCH_TILES; if (v < 0) v = 0; }
    return (u32) v;
}

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
#define SWEEP_LAUNCH(IOTA_, TICKET_) do { \
        auto kern = onesweep_kernel<MPSK_SWEEP_THREADS, MPSK_SWEEP_IPT, IOTA_, TICKET_>; \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TheSweep::SMEM); \
        if (e != cudaSuccess) return (int) e; \
        kern<<<(unsigned) ntiles, MPSK_SWEEP_THREADS, TheSweep::SMEM, stream>>>( \
            (const u64 *) kin, vin, (u64 *) kout, vout, (u32) n, (u32) shift, bins, lb, ticket, sweep_prefetch_index_tiles()); } while (0)
    const bool tickets = sweep_ticket_tiles() != 0;
    if (vin == NULL) { if (tickets) SWEEP_LAUNCH(true, true); else SWEEP_LAUNCH(true, false); }
    else { if (tickets) SWEEP_LAUNCH(false, true); else SWEEP_LAUNCH(false, false); }
#undef SWEEP_LAUNCH
    CUDA_LAUNCH_CHECK();
    return 0;
}
