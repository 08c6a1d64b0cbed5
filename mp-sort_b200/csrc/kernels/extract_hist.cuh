/* kernels/extract_hist.cuh -- K1: key extraction + digit histograms (index mode and record mode).
 * Part of the single translation unit mpsort_kernels.cu (included there, in order). */
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
                u32 * __restrict__ hist, unsigned long long * __restrict__ diff, const u64 * __restrict__ ref)
{
    __shared__ u32 sh[NH * 256];
    for (u32 t = threadIdx.x; t < NH * 256; t += blockDim.x) sh[t] = 0;
    __syncthreads();
    const u64 k0 = ref[koff] ^ flip;
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
        uint32_t d0, uint32_t nh, uint32_t * hist, uint64_t * diff, const void * ref, mpsk_stream_t stream)
{
    if (n == 0) return 0;
    if ((elsize != 8 && elsize != 16) || (nh != 4 && nh != 8) || d0 + nh > 8) return (int) cudaErrorInvalidValue;
    if (!ref) ref = recs;
    const int threads = 512;
    size_t blocks = (n + (size_t) threads * EXTRACT_BATCH - 1) / ((size_t) threads * EXTRACT_BATCH);
    const size_t maxb = (size_t) num_sms() * 8;
    if (blocks > maxb) blocks = maxb;
    const u32 W = (u32) (elsize / 8), koff = (key_in_high && elsize == 16) ? 1u : 0u;
    if (nh == 4)
        rec_hist_kernel<4><<<(unsigned) blocks, threads, 0, (cudaStream_t) stream>>>(
            (const u64 *) recs, W, koff, n, (u64) flip, d0, hist, (unsigned long long *) diff, (const u64 *) ref);
    else
        rec_hist_kernel<8><<<(unsigned) blocks, threads, 0, (cudaStream_t) stream>>>(
            (const u64 *) recs, W, koff, n, (u64) flip, d0, hist, (unsigned long long *) diff, (const u64 *) ref);
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
