/* kernels/p2p_exchange.cuh -- K6: record exchange by peer stores over NVLink; fused gather + exchange candidate.
 * Part of the single translation unit mpsort_kernels.cu (included there, in order). */
/* ========================================================================= */
/* K6: record exchange by peer stores over NVLink                             */
/* ========================================================================= */
/*
 * Replaces MPI_Alltoallv / the sparse Isend-Irecv variant (mp-mpiu.c:69-236) inside
 * one box: every rank's receive buffer is mapped into all peers (CUDA IPC), and ONE
 * kernel per rank copies each destination's contiguous slice of the sorted records
 * straight into that peer's buffer with 16-byte stores (st.global on peer addresses
 * go over NVLink 5 / NVSwitch). Zero-length pairs cost nothing (the sparse variant).
 * CTAs are dealt to segments in proportion to their bytes.
 */
#define MPSK_P2P_MAX_SEGS 64
struct P2PPlan {
    const unsigned char * src[MPSK_P2P_MAX_SEGS];
    unsigned char * dst[MPSK_P2P_MAX_SEGS];
    unsigned long long bytes[MPSK_P2P_MAX_SEGS];
    u32 cta_begin[MPSK_P2P_MAX_SEGS + 1];     /* CTAs [cta_begin[k], cta_begin[k+1]) serve segment k */
    int nseg;
};

template <typename V>
__global__ void __launch_bounds__(512)
p2p_copy_kernel(P2PPlan plan)
{
    int k = 0;
    while (k + 1 < plan.nseg && blockIdx.x >= plan.cta_begin[k + 1]) k++;
    const u32 ncta = plan.cta_begin[k + 1] - plan.cta_begin[k];
    const u32 cta = blockIdx.x - plan.cta_begin[k];
    const V * __restrict__ src = (const V *) plan.src[k];
    V * __restrict__ dst = (V *) plan.dst[k];
    const size_t nv = plan.bytes[k] / sizeof(V);
    constexpr int U = 4;
    const size_t stride = (size_t) ncta * blockDim.x * U;
    for (size_t i0 = ((size_t) cta * blockDim.x) * U + threadIdx.x; i0 < nv; i0 += stride) {
        V v[U];
#pragma unroll
        for (int u = 0; u < U; u++) { const size_t i = i0 + (size_t) u * blockDim.x; if (i < nv) v[u] = src[i]; }
#pragma unroll
        for (int u = 0; u < U; u++) { const size_t i = i0 + (size_t) u * blockDim.x; if (i < nv) dst[i] = v[u]; }
    }
    __threadfence_system();        /* peer stores performed before the kernel is seen as done */
}

/* the same copy with the whole grid on ONE segment at a time, segments in the caller's
 * order (rotated by rank: the classic shifted all-to-all schedule, every GPU sends to
 * one peer and receives from one peer at a time) */
template <typename V>
__global__ void __launch_bounds__(512)
p2p_copy_seq_kernel(P2PPlan plan)
{
    constexpr int U = 4;
    const size_t stride = (size_t) gridDim.x * blockDim.x * U;
    for (int k = 0; k < plan.nseg; k++) {
        const V * __restrict__ src = (const V *) plan.src[k];
        V * __restrict__ dst = (V *) plan.dst[k];
        const size_t nv = plan.bytes[k] / sizeof(V);
        for (size_t i0 = ((size_t) blockIdx.x * blockDim.x) * U + threadIdx.x; i0 < nv; i0 += stride) {
            V v[U];
#pragma unroll
            for (int u = 0; u < U; u++) { const size_t i = i0 + (size_t) u * blockDim.x; if (i < nv) v[u] = src[i]; }
#pragma unroll
            for (int u = 0; u < U; u++) { const size_t i = i0 + (size_t) u * blockDim.x; if (i < nv) dst[i] = v[u]; }
        }
    }
    __threadfence_system();
}

extern "C" int mpsk_p2p_alltoallv(const void * const * src, void * const * dst, const uint64_t * bytes,
        const unsigned char * remote, int nseg, mpsk_stream_t stream)
{
    /* CTAs are dealt by bytes; weighting remote bytes higher did not help (profiles/r01_p2p_exchange.log) */
    static int wremote = -1, cta_mult = -1, sequential = -1;
    if (sequential < 0) sequential = getenv("MPSORT_P2P_SEQUENTIAL") ? 1 : 0;
    if (wremote < 0) { const char * e = getenv("MPSORT_P2P_REMOTE_WEIGHT"); wremote = e ? atoi(e) : 1; }
    if (cta_mult < 0) { const char * e = getenv("MPSORT_P2P_CTAS_PER_SM"); cta_mult = e ? atoi(e) : (nseg > 2 ? 1 : 4); }
    double weight[MPSK_P2P_MAX_SEGS];
    if (nseg > MPSK_P2P_MAX_SEGS) return (int) cudaErrorInvalidValue;
    P2PPlan plan;
    unsigned long long total = 0;
    uintptr_t align = 0;
    int n = 0;
    for (int k = 0; k < nseg; k++) {
        if (bytes[k] == 0) continue;
        plan.src[n] = (const unsigned char *) src[k];
        plan.dst[n] = (unsigned char *) dst[k];
        plan.bytes[n] = bytes[k];
        weight[n] = (double) bytes[k] * (remote[k] ? wremote : 1);
        total += (unsigned long long) weight[n];
        align |= (uintptr_t) src[k] | (uintptr_t) dst[k] | (uintptr_t) bytes[k];
        n++;
    }
    if (n == 0) return 0;
    plan.nseg = n;
    const u32 G = (u32) num_sms() * (u32) cta_mult;
    u32 acc = 0;
    for (int k = 0; k < n; k++) {
        u32 share = (u32) ((double) G * weight[k] / (double) total);
        if (share < 1) share = 1;
        plan.cta_begin[k] = acc;
        acc += share;
    }
    plan.cta_begin[n] = acc;
    cudaStream_t st = (cudaStream_t) stream;
    if (sequential) {
        if ((align & 15) == 0) p2p_copy_seq_kernel<uint4><<<G, 512, 0, st>>>(plan);
        else if ((align & 7) == 0) p2p_copy_seq_kernel<u64><<<G, 512, 0, st>>>(plan);
        else if ((align & 3) == 0) p2p_copy_seq_kernel<u32><<<G, 512, 0, st>>>(plan);
        else p2p_copy_seq_kernel<unsigned char><<<G, 512, 0, st>>>(plan);
        CUDA_LAUNCH_CHECK();
        return 0;
    }
    if ((align & 15) == 0) p2p_copy_kernel<uint4><<<acc, 512, 0, st>>>(plan);
    else if ((align & 7) == 0) p2p_copy_kernel<u64><<<acc, 512, 0, st>>>(plan);
    else if ((align & 3) == 0) p2p_copy_kernel<u32><<<acc, 512, 0, st>>>(plan);
    else p2p_copy_kernel<unsigned char><<<acc, 512, 0, st>>>(plan);
    CUDA_LAUNCH_CHECK();
    return 0;
}

/*
 * CANDIDATE, off by default (MPSORT_FUSED_PACK=1) and not yet run on a GPU: pack and exchange of
 * index mode in ONE kernel. gather_records_kernel writes the records in sorted order into a send
 * buffer and the slices then travel by DMA; here record i of the sorted order is read from
 * base[idx[i]] and stored straight into its destination rank's receive buffer (peer memory mapped
 * with CUDA IPC, st.global over NVLink; the own slice into the local receive buffer): one pass over
 * the records instead of two, no send buffer. Segments as in p2p_copy_kernel.
 */
struct P2PGatherPlan {
    const u32 * idx[MPSK_P2P_MAX_SEGS];          /* sorted-order source positions of the segment's records */
    unsigned char * dst[MPSK_P2P_MAX_SEGS];      /* where the segment's first record lands */
    unsigned long long nrec[MPSK_P2P_MAX_SEGS];
    u32 cta_begin[MPSK_P2P_MAX_SEGS + 1];
    int nseg;
};

template <typename V>
__global__ void __launch_bounds__(512)
p2p_gather_kernel(const V * __restrict__ base, P2PGatherPlan plan, u32 lpr /* V pieces per record */)
{
    int k = 0;
    while (k + 1 < plan.nseg && blockIdx.x >= plan.cta_begin[k + 1]) k++;
    const u32 ncta = plan.cta_begin[k + 1] - plan.cta_begin[k];
    const u32 cta = blockIdx.x - plan.cta_begin[k];
    const u32 * __restrict__ idx = plan.idx[k];
    V * __restrict__ dst = (V *) plan.dst[k];
    const size_t nv = (size_t) plan.nrec[k] * lpr;
    constexpr int U = 4;
    const size_t stride = (size_t) ncta * blockDim.x * U;
    for (size_t i0 = ((size_t) cta * blockDim.x) * U + threadIdx.x; i0 < nv; i0 += stride) {
        size_t src[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t i = i0 + (size_t) u * blockDim.x;
            src[u] = 0;
            if (i < nv) {
                const size_t rec = i / lpr;
                src[u] = (size_t) idx[rec] * lpr + (i - rec * lpr);
            }
        }
        V v[U];
#pragma unroll
        for (int u = 0; u < U; u++) { const size_t i = i0 + (size_t) u * blockDim.x; if (i < nv) v[u] = base[src[u]]; }
#pragma unroll
        for (int u = 0; u < U; u++) { const size_t i = i0 + (size_t) u * blockDim.x; if (i < nv) dst[i] = v[u]; }
    }
    __threadfence_system();        /* peer stores performed before the kernel is seen as done */
}

template <typename V>
static int launch_p2p_gather(const void * base, const P2PGatherPlan & plan, u32 grid, size_t elsize, cudaStream_t st)
{
    p2p_gather_kernel<V><<<grid, 512, 0, st>>>((const V *) base, plan, (u32) (elsize / sizeof(V)));
    CUDA_LAUNCH_CHECK();
    return 0;
}

extern "C" int mpsk_p2p_gather_alltoallv(const void * base, const uint32_t * const * idx, void * const * dst,
        const uint64_t * nrec, size_t elsize, int nseg, mpsk_stream_t stream)
{
    if (nseg > MPSK_P2P_MAX_SEGS || elsize == 0) return (int) cudaErrorInvalidValue;
    static int cta_mult = -1;
    if (cta_mult < 0) { const char * e = getenv("MPSORT_P2P_CTAS_PER_SM"); cta_mult = e ? atoi(e) : 2; if (cta_mult < 1) cta_mult = 1; }
    P2PGatherPlan plan;
    unsigned long long total = 0;
    uintptr_t align = (uintptr_t) base | (uintptr_t) elsize;
    int n = 0;
    for (int k = 0; k < nseg; k++) {
        if (nrec[k] == 0) continue;
        plan.idx[n] = idx[k];
        plan.dst[n] = (unsigned char *) dst[k];
        plan.nrec[n] = nrec[k];
        total += nrec[k];
        align |= (uintptr_t) dst[k];
        n++;
    }
    if (n == 0) return 0;
    plan.nseg = n;
    const u32 G = (u32) num_sms() * (u32) cta_mult;
    u32 acc = 0;
    for (int k = 0; k < n; k++) {
        u32 share = (u32) ((double) G * (double) plan.nrec[k] / (double) total);
        if (share < 1) share = 1;
        plan.cta_begin[k] = acc;
        acc += share;
    }
    plan.cta_begin[n] = acc;
    cudaStream_t st = (cudaStream_t) stream;
    if ((align & 15) == 0) return launch_p2p_gather<uint4>(base, plan, acc, elsize, st);
    if ((align & 7) == 0) return launch_p2p_gather<u64>(base, plan, acc, elsize, st);
    if ((align & 3) == 0) return launch_p2p_gather<u32>(base, plan, acc, elsize, st);
    if ((align & 1) == 0) return launch_p2p_gather<unsigned short>(base, plan, acc, elsize, st);
    return launch_p2p_gather<unsigned char>(base, plan, acc, elsize, st);
}
