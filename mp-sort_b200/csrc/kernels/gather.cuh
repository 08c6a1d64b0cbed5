/* kernels/gather.cuh -- K3: payload gathers.
 * Part of the single translation unit mpsort_kernels.cu (included there, in order). */
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
