/* kernels/support.cuh -- bench / test support: synthetic records, order check.
 * Part of the single translation unit mpsort_kernels.cu (included there, in order). */
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
