/*
 * mpsort_util.c -- bench / test support (see include/mpsort_util.h). Not on the
 * product path.
 */
#include <string.h>

#include "mpsort_internal.h"
#include "mpsort_util.h"

int mpsort_util_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int mpsort_util_device_pci_bus_id(int device, char * buf, int len)
{
    if (cudaDeviceGetPCIBusId(buf, len, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    return 0;
}

void * mpsort_util_dev_malloc(int device, size_t nbytes)
{
    void * p = NULL;
    CUDA_OK(NULL, cudaSetDevice(device));
    CUDA_OK(NULL, cudaMalloc(&p, nbytes ? nbytes : 256));
    return p;
}

void mpsort_util_dev_free(int device, void * ptr)
{
    CUDA_OK(NULL, cudaSetDevice(device));
    CUDA_OK(NULL, cudaFree(ptr));
}

void * mpsort_util_host_malloc_pinned(size_t nbytes)
{
    /* NULL (not fatal) when the box cannot pin that much: the caller may fall back to
     * pageable memory */
    void * p = NULL;
    if (cudaMallocHost(&p, nbytes ? nbytes : 256) != cudaSuccess) {
        cudaGetLastError();
        return NULL;
    }
    return p;
}

void mpsort_util_host_free_pinned(void * ptr)
{
    if (ptr) CUDA_OK(NULL, cudaFreeHost(ptr));
}

void mpsort_util_memcpy(int device, void * dst, const void * src, size_t nbytes)
{
    CUDA_OK(NULL, cudaSetDevice(device));
    if (nbytes) CUDA_OK(NULL, cudaMemcpy(dst, src, nbytes, cudaMemcpyDefault));
}

void mpsort_util_dev_memset(int device, void * dst, int value, size_t nbytes)
{
    CUDA_OK(NULL, cudaSetDevice(device));
    if (nbytes) CUDA_OK(NULL, cudaMemset(dst, value, nbytes));
}

void mpsort_util_generate(mpsort_comm_t c, void * dst, size_t n, size_t elsize, int kind, uint64_t seed)
{
    CUDA_OK(c, cudaSetDevice(c->device));
    KERN_OK(c, mpsk_generate(dst, n, elsize, kind, seed, (uint64_t) c->rank, (uint64_t) c->size, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
}

void mpsort_util_generate_as(mpsort_comm_t c, void * dst, size_t n, size_t elsize, int kind, uint64_t seed,
        uint64_t rank, uint64_t nranks)
{
    CUDA_OK(c, cudaSetDevice(c->device));
    KERN_OK(c, mpsk_generate(dst, n, elsize, kind, seed, rank, nranks, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
}

void mpsort_util_multiset_hash(mpsort_comm_t c, const void * base, size_t n, size_t elsize, uint64_t * out2)
{
    uint64_t h[2] = { 0, 0 };
    struct cudaPointerAttributes a;
    CUDA_OK(c, cudaSetDevice(c->device));
    int on_dev = 0;
    if (cudaPointerGetAttributes(&a, base) == cudaSuccess)
        on_dev = (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged);
    else cudaGetLastError();
    const void * src = base;
    if (!on_dev && n) {
        void * tmp = mps_arena_get(c, MPS_S_DIN, n * elsize);
        CUDA_OK(c, cudaMemcpyAsync(tmp, base, n * elsize, cudaMemcpyHostToDevice, c->stream));
        src = tmp;
    }
    uint64_t * d = (uint64_t *) mps_arena_get(c, MPS_S_MISC, 256);
    CUDA_OK(c, cudaMemsetAsync(d, 0, 2 * sizeof(uint64_t), c->stream));
    KERN_OK(c, mpsk_multiset_hash(src, n, elsize, d, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    out2[0] = h[0]; out2[1] = h[1];
}

uint64_t mpsort_util_check_sorted(mpsort_comm_t c, const void * base, size_t n, size_t elsize,
        const struct mpsort_radix_desc * desc, int check_ties, size_t tie_offset, uint64_t * firstlast)
{
    const uint32_t nw = (uint32_t) (((size_t) desc->width * desc->nwords + 7) / 8);
    uint64_t h[1 + 2 * MPS_MAX_KEY_WORDS];
    CUDA_OK(c, cudaSetDevice(c->device));
    uint64_t * d = (uint64_t *) mps_arena_get(c, MPS_S_MISC, sizeof(h));
    CUDA_OK(c, cudaMemsetAsync(d, 0, sizeof(h), c->stream));
    KERN_OK(c, mpsk_check_sorted(base, n, elsize, desc->offset, desc->width, desc->nwords,
                                 desc->is_signed, check_ties, tie_offset, d, d + 1, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    if (firstlast) memcpy(firstlast, h + 1, 2 * (size_t) nw * sizeof(uint64_t));
    return h[0];
}

uint64_t mpsort_util_checksum(mpsort_comm_t c, const void * base, size_t nbytes)
{
    uint64_t h = 0;
    struct cudaPointerAttributes a;
    CUDA_OK(c, cudaSetDevice(c->device));
    int on_dev = 0;
    if (cudaPointerGetAttributes(&a, base) == cudaSuccess)
        on_dev = (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged);
    else cudaGetLastError();
    const void * src = base;
    if (!on_dev && nbytes) {
        void * tmp = mps_arena_get(c, MPS_S_DIN, nbytes);
        CUDA_OK(c, cudaMemcpyAsync(tmp, base, nbytes, cudaMemcpyHostToDevice, c->stream));
        src = tmp;
    }
    uint64_t * d = (uint64_t *) mps_arena_get(c, MPS_S_MISC, 256);
    CUDA_OK(c, cudaMemsetAsync(d, 0, sizeof(uint64_t), c->stream));
    KERN_OK(c, mpsk_checksum(src, nbytes, d, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return h;
}

void * mpsort_util_event_create(mpsort_comm_t c)
{
    cudaEvent_t e;
    CUDA_OK(c, cudaSetDevice(c->device));
    CUDA_OK(c, cudaEventCreate(&e));
    return (void *) e;
}

void mpsort_util_event_record(mpsort_comm_t c, void * event)
{
    CUDA_OK(c, cudaSetDevice(c->device));
    CUDA_OK(c, cudaEventRecord((cudaEvent_t) event, c->stream));
}

double mpsort_util_event_elapsed_ms(mpsort_comm_t c, void * start, void * stop)
{
    float ms = 0;
    CUDA_OK(c, cudaSetDevice(c->device));
    CUDA_OK(c, cudaEventSynchronize((cudaEvent_t) stop));
    CUDA_OK(c, cudaEventElapsedTime(&ms, (cudaEvent_t) start, (cudaEvent_t) stop));
    return (double) ms;
}

void mpsort_util_event_destroy(void * event)
{
    cudaEventDestroy((cudaEvent_t) event);
}

void mpsort_util_stream_sync(mpsort_comm_t c)
{
    CUDA_OK(c, cudaSetDevice(c->device));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
}

void mpsort_util_flush_l2(mpsort_comm_t c)
{
    const size_t nbytes = (size_t) 256 << 20;
    CUDA_OK(c, cudaSetDevice(c->device));
    void * p = mps_arena_get(c, MPS_S_STAGE2, nbytes);
    CUDA_OK(c, cudaMemsetAsync(p, 0x5a, nbytes, c->stream));
}

extern uint64_t mpsk_launch_count(int reset);
uint64_t mpsort_util_launch_count(int reset)
{
    return mpsk_launch_count(reset);
}

void mpsort_util_mem_info(int device, size_t * free_bytes, size_t * total_bytes)
{
    CUDA_OK(NULL, cudaSetDevice(device));
    CUDA_OK(NULL, cudaMemGetInfo(free_bytes, total_bytes));
}

static const char * kclass_names[MPS_NKCLASS] = {
    "extract_hist", "onesweep_pass", "onesweep_pass_rec16", "gather_keys", "gather_records", "splitter", "checksum", "exchange", "merge_runs", "hybrid_fixup"
};

void mpsort_util_kernel_timing(mpsort_comm_t c, int on)
{
    int i;
    CUDA_OK(c, cudaSetDevice(c->device));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    mps_kt_collect(c);
    for (i = 0; i < MPS_NKCLASS; i++) { c->kt.ms[i] = 0; c->kt.launches[i] = 0; }
    c->kt.on = on;
}

int mpsort_util_kernel_times(mpsort_comm_t c, const char ** names, double * ms, uint64_t * launches, int max)
{
    int i;
    for (i = 0; i < MPS_NKCLASS && i < max; i++) {
        if (names) names[i] = kclass_names[i];
        if (ms) ms[i] = c->kt.ms[i];
        if (launches) launches[i] = c->kt.launches[i];
    }
    return MPS_NKCLASS;
}
