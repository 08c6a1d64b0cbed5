/*
 * mock_device.c -- TEST INFRASTRUCTURE ONLY. A stand-in for the GPU side of libmpsort-b200.so so
 * that the product's C HOST code (mpsort_host.c, mpsort_comm.c, mpsort_layout.c, mpsort_util.c --
 * compiled unchanged) can be driven on a CPU-only box: tests/test_hostflow_mock.py links those
 * files against this one into tests/native/_build/libmpsort-hostmock.so and checks the host
 * orchestration (phases, layout, exchange parts, gather path, callback entry points, arena use)
 * against the oracle with rank threads.
 *
 * What is mocked: (1) the CUDA runtime calls the host code makes -- "device memory" is malloc,
 * every stream operation runs synchronously, events carry host wall time; (2) the kernel ABI of
 * mpsort_kernels.h -- each entry point is restated as a plain loop with the SAME contract
 * (what it reads, what it writes, stability), not the same algorithm; (3) NCCL -- absent: only
 * the in-process transport (rank threads) works here.
 *
 * What this can NOT show: anything about the real kernels, stream ordering, races, peer memory,
 * NCCL. It is never built by the product Makefile, never shipped, and the product library has no
 * CPU path: it aborts without a CUDA device.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <cuda_runtime_api.h>
#include <nccl.h>

#include "mpsort_kernels.h"
#include "synth.h"       /* oracle/: the generator restated on the CPU (this file is a checker, not the product) */

/* ------------------------------------------------------------------------- */
/* CUDA runtime                                                               */

#define MAXALLOC 4096
static struct { char * p; size_t n; } g_alloc[MAXALLOC];
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

cudaError_t cudaMalloc(void ** p, size_t n)
{
    int i;
    *p = malloc(n ? n : 1);
    if (!*p) return cudaErrorMemoryAllocation;
    pthread_mutex_lock(&g_lock);
    for (i = 0; i < MAXALLOC; i++) if (!g_alloc[i].p) { g_alloc[i].p = (char *) *p; g_alloc[i].n = n ? n : 1; break; }
    pthread_mutex_unlock(&g_lock);
    return i < MAXALLOC ? cudaSuccess : cudaErrorMemoryAllocation;
}

cudaError_t cudaFree(void * p)
{
    int i;
    if (!p) return cudaSuccess;
    pthread_mutex_lock(&g_lock);
    for (i = 0; i < MAXALLOC; i++) if (g_alloc[i].p == (char *) p) { g_alloc[i].p = NULL; break; }
    pthread_mutex_unlock(&g_lock);
    free(p);
    return cudaSuccess;
}

cudaError_t cudaPointerGetAttributes(struct cudaPointerAttributes * a, const void * ptr)
{
    int i, dev = 0;
    pthread_mutex_lock(&g_lock);
    for (i = 0; i < MAXALLOC; i++)
        if (g_alloc[i].p && (const char *) ptr >= g_alloc[i].p && (const char *) ptr < g_alloc[i].p + g_alloc[i].n) { dev = 1; break; }
    pthread_mutex_unlock(&g_lock);
    memset(a, 0, sizeof(*a));
    a->type = dev ? cudaMemoryTypeDevice : cudaMemoryTypeUnregistered;
    return cudaSuccess;
}

cudaError_t cudaMallocHost(void ** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void * p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void * d, const void * s, size_t n, enum cudaMemcpyKind k) { (void) k; if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void * d, const void * s, size_t n, enum cudaMemcpyKind k, cudaStream_t st) { (void) k; (void) st; if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void * d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void * d, int v, size_t n, cudaStream_t st) { (void) st; if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t * s, unsigned f) { (void) f; *s = (cudaStream_t) malloc(8); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t s) { (void) s; return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned f) { (void) s; (void) e; (void) f; return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t * e) { *e = (cudaEvent_t) malloc(8); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t * e, unsigned f) { (void) f; *e = (cudaEvent_t) malloc(8); return cudaSuccess; }
static double now_ms(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return 1e3 * (double) t.tv_sec + 1e-6 * (double) t.tv_nsec; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) { (void) s; *(double *) e = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t e) { (void) e; return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float * ms, cudaEvent_t a, cudaEvent_t b)
{
    /* host wall time between the two records (everything is synchronous here), never zero */
    const double d = *(double *) b - *(double *) a;
    *ms = (float) (d > 1e-3 ? d : 1e-3);
    return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { (void) d; return cudaSuccess; }
cudaError_t cudaGetDevice(int * d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int * n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char * cudaGetErrorString(cudaError_t e) { (void) e; return "mock device error"; }
cudaError_t cudaDeviceSetLimit(enum cudaLimit l, size_t v) { (void) l; (void) v; return cudaSuccess; }
cudaError_t cudaDeviceCanAccessPeer(int * can, int a, int b) { (void) a; (void) b; *can = 1; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int d, unsigned f) { (void) d; (void) f; return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t * h, void * p) { (void) h; (void) p; return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void ** p, cudaIpcMemHandle_t h, unsigned f) { (void) p; (void) h; (void) f; return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void * p) { (void) p; return cudaSuccess; }
cudaError_t cudaMemGetInfo(size_t * fr, size_t * tot) { *fr = (size_t) 1 << 34; *tot = (size_t) 1 << 35; return cudaSuccess; }

/* ------------------------------------------------------------------------- */
/* NCCL: absent                                                               */

ncclResult_t ncclGetUniqueId(ncclUniqueId * id) { (void) id; return ncclInternalError; }
ncclResult_t ncclCommInitRank(ncclComm_t * c, int n, ncclUniqueId id, int r) { (void) c; (void) n; (void) id; (void) r; return ncclInternalError; }
ncclResult_t ncclCommDestroy(ncclComm_t c) { (void) c; return ncclSuccess; }
ncclResult_t ncclCommAbort(ncclComm_t c) { (void) c; return ncclSuccess; }
const char * ncclGetErrorString(ncclResult_t r) { (void) r; return "no NCCL in the mock"; }
ncclResult_t ncclGroupStart(void) { return ncclInternalError; }
ncclResult_t ncclGroupEnd(void) { return ncclInternalError; }
ncclResult_t ncclAllReduce(const void * s, void * r, size_t n, ncclDataType_t t, ncclRedOp_t o, ncclComm_t c, cudaStream_t st)
{ (void) s; (void) r; (void) n; (void) t; (void) o; (void) c; (void) st; return ncclInternalError; }
ncclResult_t ncclAllGather(const void * s, void * r, size_t n, ncclDataType_t t, ncclComm_t c, cudaStream_t st)
{ (void) s; (void) r; (void) n; (void) t; (void) c; (void) st; return ncclInternalError; }
ncclResult_t ncclBroadcast(const void * s, void * r, size_t n, ncclDataType_t t, int root, ncclComm_t c, cudaStream_t st)
{ (void) s; (void) r; (void) n; (void) t; (void) root; (void) c; (void) st; return ncclInternalError; }
ncclResult_t ncclSend(const void * s, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t st)
{ (void) s; (void) n; (void) t; (void) peer; (void) c; (void) st; return ncclInternalError; }
ncclResult_t ncclRecv(void * r, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t st)
{ (void) r; (void) n; (void) t; (void) peer; (void) c; (void) st; return ncclInternalError; }

/* ------------------------------------------------------------------------- */
/* kernel ABI (mpsort_kernels.h), restated as loops                           */

static unsigned long long g_launches;
#define LAUNCHED() __atomic_fetch_add(&g_launches, 1ULL, __ATOMIC_RELAXED)
#define NOT_MOCKED (LAUNCHED(), (int) cudaErrorNotSupported)

uint64_t mpsk_launch_count(int reset)
{
    const unsigned long long v = __atomic_load_n(&g_launches, __ATOMIC_RELAXED);
    if (reset) __atomic_store_n(&g_launches, 0ULL, __ATOMIC_RELAXED);
    return v;
}

static uint64_t load_le(const unsigned char * p, uint32_t width)
{
    uint64_t v = 0;
    uint32_t b;
    for (b = 0; b < width; b++) v |= ((uint64_t) p[b]) << (8 * b);
    return v;
}

/* packed 64-bit word g of the key: key bytes [8g, 8g+8) of the little-endian string of the
 * (sign-flipped) key words */
static uint64_t pack_word(const unsigned char * rec, size_t offset, uint32_t width, uint32_t nwords, int is_signed, uint32_t g)
{
    const uint32_t per = 8 / width, first = g * per;
    uint64_t out = 0;
    uint32_t k;
    for (k = 0; k < per; k++) {
        const uint32_t wi = first + k;
        if (wi >= nwords) break;
        uint64_t v = load_le(rec + offset + (size_t) wi * width, width);
        if (is_signed) v ^= 1ULL << (8 * width - 1);
        out |= v << (8 * width * k);
    }
    return out;
}

int mpsk_extract_keys(const void * base, size_t n, size_t elsize, size_t offset, uint32_t width, uint32_t nwords,
        int is_signed, uint32_t g, uint64_t sub, uint64_t * kout, uint32_t * hist, uint64_t * minmax, mpsk_stream_t stream)
{
    size_t i;
    int d;
    (void) stream;
    if (n == 0) return 0;
    LAUNCHED();
    const int inplace = (base == (const void *) kout && elsize == 8);
    for (i = 0; i < n; i++) {
        uint64_t k = inplace ? kout[i] : pack_word((const unsigned char *) base + i * elsize, offset, width, nwords, is_signed, g);
        k -= sub;
        if (kout) kout[i] = k;
        for (d = 0; d < 8; d++) hist[d * 256 + ((k >> (8 * d)) & 255u)]++;
        if (minmax) { if (k < minmax[0]) minmax[0] = k; if (k > minmax[1]) minmax[1] = k; }
    }
    return 0;
}

static uint64_t rec_key(const void * recs, size_t i, size_t elsize, int khi, uint64_t flip)
{
    uint64_t k;
    memcpy(&k, (const char *) recs + i * elsize + ((khi && elsize == 16) ? 8 : 0), 8);
    return k ^ flip;
}

int mpsk_rec_histograms(const void * recs, size_t n, size_t elsize, int khi, uint64_t flip, uint32_t d0, uint32_t nh,
        uint32_t * hist, uint64_t * diff, mpsk_stream_t stream)
{
    size_t i;
    uint32_t q;
    (void) stream;
    if (n == 0) return 0;
    if ((elsize != 8 && elsize != 16) || (nh != 4 && nh != 8) || d0 + nh > 8) return (int) cudaErrorInvalidValue;
    LAUNCHED();
    const uint64_t k0 = rec_key(recs, 0, elsize, khi, flip);
    uint64_t acc = 0;
    for (i = 0; i < n; i++) {
        const uint64_t k = rec_key(recs, i, elsize, khi, flip);
        acc |= k ^ k0;
        for (q = 0; q < nh; q++) hist[(d0 + q) * 256 + ((k >> (8 * (d0 + q))) & 255u)]++;
    }
    if (diff) *diff |= acc;
    return 0;
}

int mpsk_rec_sample_diff(const void * recs, size_t n, size_t elsize, int khi, uint32_t s, uint64_t * diff, mpsk_stream_t stream)
{
    uint32_t i;
    (void) stream;
    if (n == 0 || s == 0) return 0;
    LAUNCHED();
    const uint64_t k0 = rec_key(recs, 0, elsize, khi, 0);
    for (i = 0; i < s; i++) {
        const size_t pos = (size_t) (((unsigned __int128) i * n) / s);
        *diff |= rec_key(recs, pos, elsize, khi, 0) ^ k0;
    }
    return 0;
}

int mpsk_scan_histograms(const uint32_t * hist, uint32_t * bins, int nhist, mpsk_stream_t stream)
{
    int h, b;
    (void) stream;
    if (nhist <= 0) return 0;
    LAUNCHED();
    for (h = 0; h < nhist; h++) {
        uint32_t run = 0;
        for (b = 0; b < 256; b++) { const uint32_t c = hist[h * 256 + b]; bins[h * 256 + b] = run; run += c; }
    }
    return 0;
}

size_t mpsk_onesweep_tile_items(void) { return 6144; }
size_t mpsk_onesweep_scratch_bytes(size_t n) { return 4096 + n / 8; }

int mpsk_onesweep_pass(const uint64_t * kin, const uint32_t * vin, uint64_t * kout, uint32_t * vout, size_t n, int shift,
        const uint32_t * bins, void * scratch, mpsk_stream_t stream)
{
    uint32_t next[256];
    size_t i;
    (void) scratch; (void) stream;
    if (n == 0) return 0;
    if (n > MPSK_MAX_ITEMS) return (int) cudaErrorInvalidValue;
    LAUNCHED();
    memcpy(next, bins, sizeof(next));
    for (i = 0; i < n; i++) {
        const uint32_t at = next[(kin[i] >> shift) & 255u]++;
        if (kout) kout[at] = kin[i];
        if (vout) vout[at] = vin ? vin[i] : (uint32_t) i;
    }
    return 0;
}

int mpsk_onesweep_pass_rec(const void * in, void * out, size_t n, size_t elsize, int shift, int khi, uint64_t flip,
        const uint32_t * bins, void * scratch, mpsk_stream_t stream)
{
    uint32_t next[256];
    size_t i;
    (void) scratch; (void) stream;
    if (n == 0) return 0;
    if (elsize != 8 && elsize != 16) return (int) cudaErrorInvalidValue;
    LAUNCHED();
    memcpy(next, bins, sizeof(next));
    for (i = 0; i < n; i++) {
        const uint32_t at = next[(rec_key(in, i, elsize, khi, flip) >> shift) & 255u]++;
        memcpy((char *) out + (size_t) at * elsize, (const char *) in + i * elsize, elsize);
    }
    return 0;
}

int mpsk_fixup_rec(void * recs, size_t n, size_t elsize, int khi, uint64_t flip, uint32_t lobits,
        uint32_t * worklist, uint32_t * nwork, uint32_t cap, mpsk_stream_t stream)
{
    size_t i = 0;
    (void) stream;
    if (n == 0) return 0;
    LAUNCHED();
    const uint64_t lomask = lobits >= 64 ? ~0ULL : ((1ULL << lobits) - 1ULL);
    char * tmp = (char *) malloc(256 * elsize);
    while (i < n) {
        size_t j = i + 1, a, b;
        const uint64_t hi = lobits >= 64 ? 0 : rec_key(recs, i, elsize, khi, flip) >> lobits;
        while (j < n && (lobits >= 64 ? 0 : rec_key(recs, j, elsize, khi, flip) >> lobits) == hi) j++;
        if (j - i > 256) {
            const uint32_t slot = (*nwork)++;
            if (slot < cap) worklist[slot] = (uint32_t) i;
        } else if (j - i > 1) {
            /* stable insertion by the low part */
            memcpy(tmp, (char *) recs + i * elsize, (j - i) * elsize);
            for (a = 1; a < j - i; a++) {
                char one[16];
                memcpy(one, tmp + a * elsize, elsize);
                uint64_t ka;
                memcpy(&ka, one + ((khi && elsize == 16) ? 8 : 0), 8);
                ka = (ka ^ flip) & lomask;
                b = a;
                while (b > 0) {
                    uint64_t kb;
                    memcpy(&kb, tmp + (b - 1) * elsize + ((khi && elsize == 16) ? 8 : 0), 8);
                    kb = (kb ^ flip) & lomask;
                    if (kb <= ka) break;
                    memcpy(tmp + b * elsize, tmp + (b - 1) * elsize, elsize);
                    b--;
                }
                memcpy(tmp + b * elsize, one, elsize);
            }
            memcpy((char *) recs + i * elsize, tmp, (j - i) * elsize);
        }
        i = j;
    }
    free(tmp);
    return 0;
}

int mpsk_fixup_extents(const void * recs, size_t n, size_t elsize, int khi, uint64_t flip, uint32_t lobits,
        const uint32_t * worklist, uint32_t nwork, uint32_t * lengths, mpsk_stream_t stream)
{
    uint32_t e;
    (void) stream;
    if (nwork == 0) return 0;
    LAUNCHED();
    for (e = 0; e < nwork; e++) {
        size_t j = worklist[e];
        const uint64_t hi = rec_key(recs, j, elsize, khi, flip) >> lobits;
        while (j < n && (rec_key(recs, j, elsize, khi, flip) >> lobits) == hi) j++;
        lengths[e] = (uint32_t) (j - worklist[e]);
    }
    return 0;
}

int mpsk_sample_prefix_rec(const void * recs, size_t n, size_t elsize, uint32_t s, int khi, uint64_t flip, uint32_t lobits,
        uint64_t * out, mpsk_stream_t stream)
{
    uint32_t i;
    (void) stream;
    if (s == 0) return 0;
    LAUNCHED();
    for (i = 0; i < s; i++) out[i] = rec_key(recs, (size_t) (((unsigned __int128) i * n) / s), elsize, khi, flip) >> lobits;
    return 0;
}

int mpsk_count_equal_pairs(const uint64_t * sorted, uint32_t s, uint64_t * count, mpsk_stream_t stream)
{
    uint32_t i, first = 0;
    (void) stream;
    if (s == 0) return 0;
    LAUNCHED();
    for (i = 0; i < s; i++) {
        if (sorted[i] != sorted[first]) first = i;
        *count += i - first;
    }
    return 0;
}

int mpsk_gather_u64(const uint64_t * src, const uint32_t * idx, uint64_t * dst, size_t n, mpsk_stream_t stream)
{
    size_t i;
    (void) stream;
    if (n == 0) return 0;
    LAUNCHED();
    for (i = 0; i < n; i++) dst[i] = src[idx[i]];
    return 0;
}

int mpsk_gather_records(const void * base, const uint32_t * idx, void * out, size_t n, size_t elsize, mpsk_stream_t stream)
{
    size_t i;
    (void) stream;
    if (n == 0 || elsize == 0) return 0;
    LAUNCHED();
    for (i = 0; i < n; i++) memcpy((char *) out + i * elsize, (const char *) base + (size_t) idx[i] * elsize, elsize);
    return 0;
}

/* -1, 0, +1: key i of the view against cand[] (word nw-1 most significant) */
static int cmp_view(struct mpsk_keyview v, size_t i, const uint64_t * cand, uint32_t nw)
{
    int w;
    for (w = (int) nw - 1; w >= 0; w--) {
        uint64_t k;
        memcpy(&k, (const char *) v.base + i * v.item_stride + (size_t) w * v.word_stride, 8);
        k = (k ^ v.flip) + (w == 0 ? v.add : 0ULL);
        if (k < cand[w]) return -1;
        if (k > cand[w]) return 1;
    }
    return 0;
}

static uint64_t bound_view(struct mpsk_keyview v, size_t n, const uint64_t * cand, uint32_t nw, int upper)
{
    size_t lo = 0, hi = n;
    while (lo < hi) {
        const size_t mid = lo + ((hi - lo) >> 1);
        const int c = cmp_view(v, mid, cand, nw);
        if (upper ? (c <= 0) : (c < 0)) lo = mid + 1; else hi = mid;
    }
    return (uint64_t) lo;
}

int mpsk_splitter_count(struct mpsk_keyview view, size_t n, uint32_t nw, const uint64_t * prefix, int nsplit, int level,
        uint64_t * counts, mpsk_stream_t stream)
{
    int b;
    uint32_t d, w;
    (void) stream;
    if (nsplit <= 0) return 0;
    if (nw > 16) return (int) cudaErrorInvalidValue;
    LAUNCHED();
    const uint32_t byteidx = 8 * nw - 1 - (uint32_t) level, wi = byteidx >> 3, sh = (byteidx & 7) * 8;
    for (b = 0; b < nsplit; b++)
        for (d = 0; d < 256; d++) {
            uint64_t cand[16];
            for (w = 0; w < nw; w++) {
                uint64_t x = prefix[(size_t) b * nw + w];
                if (w < wi) x = ~0ULL;
                else if (w == wi) x |= ((uint64_t) d << sh) | ((sh == 0) ? 0ULL : ((1ULL << sh) - 1ULL));
                cand[w] = x;
            }
            counts[(size_t) b * 256 + d] = bound_view(view, n, cand, nw, 1);
        }
    return 0;
}

int mpsk_splitter_select(const uint64_t * counts, const uint64_t * target, uint64_t * prefix, uint32_t nw, int nsplit, int level,
        mpsk_stream_t stream)
{
    int b;
    uint32_t d;
    (void) stream;
    if (nsplit <= 0) return 0;
    LAUNCHED();
    const uint32_t byteidx = 8 * nw - 1 - (uint32_t) level, wi = byteidx >> 3, sh = (byteidx & 7) * 8;
    for (b = 0; b < nsplit; b++) {
        uint32_t pick = 255;
        for (d = 0; d < 256; d++) if (counts[(size_t) b * 256 + d] >= target[b]) { pick = d; break; }
        prefix[(size_t) b * nw + wi] |= ((uint64_t) pick) << sh;
    }
    return 0;
}

int mpsk_splitter_final(struct mpsk_keyview view, size_t n, uint32_t nw, const uint64_t * prefix, int nsplit, uint64_t * out,
        mpsk_stream_t stream)
{
    int b;
    (void) stream;
    if (nsplit <= 0) return 0;
    LAUNCHED();
    for (b = 0; b < nsplit; b++) {
        out[b] = bound_view(view, n, prefix + (size_t) b * nw, nw, 0);
        out[nsplit + b] = bound_view(view, n, prefix + (size_t) b * nw, nw, 1);
    }
    return 0;
}

size_t mpsk_peer_box_bytes(void) { return 4096; }
int mpsk_splitter_descent_peer(struct mpsk_keyview view, size_t n, uint32_t nw, uint64_t * prefix, const uint64_t * target,
        int nsplit, int level0, int nlevels, uint32_t me, uint32_t p, void * const * boxes, uint32_t seq, uint32_t * err,
        mpsk_stream_t stream)
{
    (void) view; (void) n; (void) nw; (void) prefix; (void) target; (void) nsplit; (void) level0; (void) nlevels;
    (void) me; (void) p; (void) boxes; (void) seq; (void) err; (void) stream;
    return NOT_MOCKED;      /* needs concurrently running kernels */
}

int mpsk_sum_u64(uint64_t * dst, const uint64_t * const * srcs, int nsrc, size_t count, mpsk_stream_t stream)
{
    size_t i;
    int k;
    (void) stream;
    if (count == 0) return 0;
    LAUNCHED();
    for (i = 0; i < count; i++) {
        uint64_t s = 0;
        for (k = 0; k < nsrc; k++) s += srcs[k][i];
        dst[i] = s;
    }
    return 0;
}

/* ---- merge of the received runs: any stable p-way merge honours the contract */
size_t mpsk_merge_tile_items(void) { return 4096; }
size_t mpsk_merge_tile_items_for(const void * recv, const void * out, size_t elsize, size_t offset, uint32_t width,
        uint32_t nwords, uint32_t p)
{
    (void) recv; (void) out; (void) elsize; (void) offset; (void) width; (void) nwords; (void) p;
    return 4096;
}

int mpsk_merge_samples(const void * recv, size_t elsize, size_t offset, uint32_t width, uint32_t nwords, int is_signed,
        uint32_t p, uint32_t S, uint32_t k, const uint32_t * rdispl, const uint32_t * sstart, uint64_t * skeys, mpsk_stream_t stream)
{
    uint32_t r, j;
    (void) k; (void) stream;
    if (sstart[p] == 0) return 0;
    LAUNCHED();
    for (r = 0; r < p; r++)
        for (j = 0; j < sstart[r + 1] - sstart[r]; j++) {
            const size_t pos = (size_t) rdispl[r] + (size_t) (j + 1) * S - 1;
            skeys[sstart[r] + j] = pack_word((const unsigned char *) recv + pos * elsize, offset, width, nwords, is_signed, 0);
        }
    return 0;
}

int mpsk_merge_runs(const void * recv, void * out, size_t elsize, size_t offset, uint32_t width, uint32_t nwords, int is_signed,
        uint32_t p, uint32_t S, uint32_t k, const uint32_t * rdispl, const uint32_t * sstart,
        const uint64_t * sorted_skeys, const uint32_t * sorted_sid, uint32_t ntiles, uint32_t * cut, uint32_t * overflow,
        mpsk_stream_t stream)
{
    uint32_t head[64], r;
    size_t o = 0;
    (void) S; (void) k; (void) sstart; (void) sorted_skeys; (void) sorted_sid; (void) ntiles; (void) cut; (void) overflow; (void) stream;
    if (p > 32) return (int) cudaErrorInvalidValue;
    LAUNCHED();
    for (r = 0; r < p; r++) head[r] = rdispl[r];
    for (;;) {
        int best = -1;
        uint64_t kb = 0;
        for (r = 0; r < p; r++) {
            if (head[r] >= rdispl[r + 1]) continue;
            const uint64_t kr = pack_word((const unsigned char *) recv + (size_t) head[r] * elsize, offset, width, nwords, is_signed, 0);
            if (best < 0 || kr < kb) { best = (int) r; kb = kr; }      /* ties: the lower run */
        }
        if (best < 0) break;
        memcpy((char *) out + o * elsize, (const char *) recv + (size_t) head[best] * elsize, elsize);
        head[best]++;
        o++;
    }
    return 0;
}

int mpsk_p2p_alltoallv(const void * const * src, void * const * dst, const uint64_t * bytes, const unsigned char * remote,
        int nseg, mpsk_stream_t stream)
{
    (void) src; (void) dst; (void) bytes; (void) remote; (void) nseg; (void) stream;
    return NOT_MOCKED;      /* peer memory: NCCL transport only */
}

int mpsk_p2p_gather_alltoallv(const void * base, const uint32_t * const * idx, void * const * dst, const uint64_t * nrec,
        size_t elsize, int nseg, mpsk_stream_t stream)
{
    (void) base; (void) idx; (void) dst; (void) nrec; (void) elsize; (void) nseg; (void) stream;
    return NOT_MOCKED;
}

int mpsk_checksum(const void * base, size_t nbytes, uint64_t * sum, mpsk_stream_t stream)
{
    size_t i;
    uint64_t s = 0;
    (void) stream;
    LAUNCHED();
    for (i = 0; i < nbytes; i++) s += (uint64_t) (int64_t) ((const signed char *) base)[i];
    *sum += s;
    return 0;
}

int mpsk_generate(void * dst, size_t n, size_t elsize, int kind, uint64_t seed, uint64_t rank, uint64_t nranks, mpsk_stream_t stream)
{
    size_t i;
    (void) stream;
    if (n == 0) return 0;
    LAUNCHED();
    for (i = 0; i < n; i++) synth_record((unsigned char *) dst + i * elsize, elsize, kind, seed, rank, nranks, n, i);
    return 0;
}

int mpsk_check_sorted(const void * base, size_t n, size_t elsize, size_t offset, uint32_t width, uint32_t nwords, int is_signed,
        int check_ties, size_t tie_offset, uint64_t * violations, uint64_t * firstlast, mpsk_stream_t stream)
{
    const uint32_t nw = (uint32_t) (((size_t) width * nwords + 7) / 8);
    const unsigned char * b = (const unsigned char *) base;
    size_t i;
    uint32_t w;
    (void) stream;
    if (n == 0) return 0;
    LAUNCHED();
    for (w = 0; w < nw; w++) {
        firstlast[w] = pack_word(b, offset, width, nwords, is_signed, w);
        firstlast[nw + w] = pack_word(b + (n - 1) * elsize, offset, width, nwords, is_signed, w);
    }
    for (i = 1; i < n; i++) {
        int c = 0, g;
        for (g = (int) nw - 1; g >= 0 && c == 0; g--) {
            const uint64_t x = pack_word(b + (i - 1) * elsize, offset, width, nwords, is_signed, (uint32_t) g);
            const uint64_t y = pack_word(b + i * elsize, offset, width, nwords, is_signed, (uint32_t) g);
            c = (x > y) - (x < y);
        }
        if (c > 0) (*violations)++;
        else if (c == 0 && check_ties && load_le(b + (i - 1) * elsize + tie_offset, 8) >= load_le(b + i * elsize + tie_offset, 8))
            (*violations)++;
    }
    return 0;
}
