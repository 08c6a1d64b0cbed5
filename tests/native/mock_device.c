/*
 * mock_device.c -- TEST INFRASTRUCTURE ONLY. A stand-in for the GPU side of libmpsort-b200.so so
 * that the product's C HOST code (mpsort_host.c, mpsort_comm.c, mpsort_layout.c, mpsort_util.c --
 * compiled unchanged) can be driven on a CPU-only box: tests/test_hostflow_mock.py links those
 * files against this one into tests/native/_build/libmpsort-hostmock.so and checks the host
 * orchestration (phases, layout, exchange parts, gather path, callback entry points, arena use)
 * against the oracle with rank threads.
 *
 * What is mocked: (1) the CUDA runtime calls the host code makes -- "device memory" is malloc,
 * every stream operation runs at once, events carry host wall time; (2) the kernel ABI of
 * mpsort_kernels.h -- each entry point is restated as a plain loop with the SAME contract
 * (what it reads, what it writes, stability), not the same algorithm; (3) NCCL and CUDA IPC --
 * between the rank THREADS of one process, every collective blocking (so both the in-process
 * transport and the NCCL transport with its mapped peer buffers run here).
 * Built on its own this file is the complete, synchronous mock. tests/support/hostmock.py builds
 * it with -DMOCK_WITH_ASYNC_LAYER: the stream-ordered entry points are then renamed
 * (mock_rename.h) and exported by mock_async.cpp instead, which can defer and interleave them
 * (MOCK_ASYNC=<seed>) so that missing synchronisation between streams shows up as wrong bytes.
 *
 * What this can NOT show: anything about the real kernels, peer-memory visibility on hardware,
 * real NCCL. It is never built by the product Makefile, never shipped, and the product library has no
 * CPU path: it aborts without a CUDA device.
 */
#ifdef MOCK_WITH_ASYNC_LAYER
#include "mock_rename.h"    /* the stream-ordered entry points become mocksync_*: mock_async.cpp decides when they run */
#endif
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

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
cudaError_t cudaDeviceGetPCIBusId(char * pciBusId, int len, int device)
{
    (void) device;
    if (len > 0) pciBusId[0] = 0;
    return cudaErrorInvalidDevice;      /* no PCI address: callers skip NUMA placement */
}

cudaError_t cudaGetDeviceCount(int * n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char * cudaGetErrorString(cudaError_t e) { (void) e; return "mock device error"; }
cudaError_t cudaDeviceSetLimit(enum cudaLimit l, size_t v) { (void) l; (void) v; return cudaSuccess; }
cudaError_t cudaDeviceCanAccessPeer(int * can, int a, int b) { (void) a; (void) b; *can = 1; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int d, unsigned f) { (void) d; (void) f; return cudaSuccess; }
/* CUDA IPC between the rank THREADS of this process (the real runtime refuses to open a handle in the
 * process that made it; here a handle is the base address of the allocation, as the real one stands for
 * the whole allocation). MOCK_NO_IPC=1: not supported, which sends the host code to its NCCL fallback. */
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t * h, void * p)
{
    int i;
    char * base = NULL;
    if (getenv("MOCK_NO_IPC")) return cudaErrorNotSupported;
    pthread_mutex_lock(&g_lock);
    for (i = 0; i < MAXALLOC; i++)
        if (g_alloc[i].p && (char *) p >= g_alloc[i].p && (char *) p < g_alloc[i].p + g_alloc[i].n) { base = g_alloc[i].p; break; }
    pthread_mutex_unlock(&g_lock);
    if (!base) return cudaErrorInvalidValue;
    memset(h, 0, sizeof(*h));
    memcpy(h, &base, sizeof(base));
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void ** p, cudaIpcMemHandle_t h, unsigned f)
{
    (void) f;
    if (getenv("MOCK_NO_IPC")) return cudaErrorNotSupported;
    memcpy(p, &h, sizeof(*p));
    return *p ? cudaSuccess : cudaErrorInvalidValue;
}
cudaError_t cudaIpcCloseMemHandle(void * p) { (void) p; return cudaSuccess; }
cudaError_t cudaMemGetInfo(size_t * fr, size_t * tot) { *fr = (size_t) 1 << 34; *tot = (size_t) 1 << 35; return cudaSuccess; }

/* ------------------------------------------------------------------------- */
/* NCCL among the THREADS of this process (real NCCL allows several ranks per process too):
 * every call blocks until all ranks of the communicator have made it -- collectives in a
 * different order on different ranks, which would corrupt or hang a real run, hang here and
 * the test's timeout reports it. Grouped calls (broadcast, send, recv) are collected per
 * thread and run at ncclGroupEnd pair by pair (a rank with nothing to post is not waited for):
 * all sends are published, then the receives copy, then the sender waits until its buffers
 * have been read.                                                                          */

#define MOCK_MAXRANKS 64
#define MOCK_MAXOPS   256
struct mock_op { int kind; const void * send; void * recv; size_t bytes; int peer; };   /* kind: 0 send, 1 recv, 2 bcast (peer = root) */
#define MOCK_FIFO 8
struct mock_fifo { unsigned long head, tail; struct { const void * ptr; size_t bytes; } msg[MOCK_FIFO]; };
struct mock_group {
    int used, size, joined, refs;
    char id[32];
    pthread_barrier_t bar;
    const void * slot[MOCK_MAXRANKS];
    void * rslot[MOCK_MAXRANKS];
    struct mock_fifo * fifo;            /* [from][to][channel]: 0 send/recv, 1 broadcast */
    pthread_mutex_t mlock;
    pthread_cond_t mcond;
};
struct ncclComm { struct mock_group * g; int rank; };
static struct mock_group g_groups[32];
static int g_uid_seq;
static __thread int t_group_depth;
static __thread int t_nops;
static __thread struct mock_op t_ops[MOCK_MAXOPS];
static __thread struct ncclComm * t_group_comm;

static size_t nccl_size(ncclDataType_t t)
{
    switch (t) {
    case ncclInt8: case ncclUint8: return 1;
    case ncclInt32: case ncclUint32: case ncclFloat32: return 4;
    case ncclInt64: case ncclUint64: case ncclFloat64: return 8;
    default: return 0;
    }
}

ncclResult_t ncclGetUniqueId(ncclUniqueId * id)
{
    memset(id, 0, sizeof(*id));
    pthread_mutex_lock(&g_lock);
    snprintf(id->internal, 32, "mock-nccl-%d-%d", (int) getpid(), ++g_uid_seq);
    pthread_mutex_unlock(&g_lock);
    return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t * c, int n, ncclUniqueId id, int r)
{
    int i;
    struct mock_group * g = NULL;
    if (n < 1 || n > MOCK_MAXRANKS || r < 0 || r >= n || id.internal[0] == 0) return ncclInvalidArgument;
    pthread_mutex_lock(&g_lock);
    for (i = 0; i < 32; i++) if (g_groups[i].used && memcmp(g_groups[i].id, id.internal, 32) == 0) { g = &g_groups[i]; break; }
    if (!g) {
        for (i = 0; i < 32; i++) if (!g_groups[i].used) { g = &g_groups[i]; break; }
        if (!g) { pthread_mutex_unlock(&g_lock); return ncclInternalError; }
        memset(g, 0, sizeof(*g));
        g->used = 1; g->size = n;
        memcpy(g->id, id.internal, 32);
        pthread_barrier_init(&g->bar, NULL, (unsigned) n);
        pthread_mutex_init(&g->mlock, NULL);
        pthread_cond_init(&g->mcond, NULL);
        g->fifo = (struct mock_fifo *) calloc((size_t) n * n * 2, sizeof(struct mock_fifo));
    }
    if (g->size != n) { pthread_mutex_unlock(&g_lock); return ncclInvalidArgument; }
    g->joined++; g->refs++;
    pthread_mutex_unlock(&g_lock);
    *c = (struct ncclComm *) malloc(sizeof(struct ncclComm));
    (*c)->g = g; (*c)->rank = r;
    pthread_barrier_wait(&g->bar);      /* like the real call: returns when every rank has joined */
    return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t c)
{
    int last;
    if (!c) return ncclSuccess;
    pthread_mutex_lock(&g_lock);
    last = (--c->g->refs == 0);
    if (last) { pthread_barrier_destroy(&c->g->bar); free(c->g->fifo); c->g->used = 0; }
    pthread_mutex_unlock(&g_lock);
    free(c);
    return ncclSuccess;
}
ncclResult_t ncclCommAbort(ncclComm_t c) { (void) c; return ncclSuccess; }
const char * ncclGetErrorString(ncclResult_t r) { (void) r; return "mock NCCL error"; }

ncclResult_t ncclAllReduce(const void * s, void * r, size_t n, ncclDataType_t t, ncclRedOp_t o, ncclComm_t c, cudaStream_t st)
{
    struct mock_group * g = c->g;
    const size_t w = nccl_size(t);
    size_t i;
    int k;
    (void) st;
    if (o != ncclSum || (w != 4 && w != 8) || t_group_depth) return ncclInvalidUsage;
    g->slot[c->rank] = s;
    pthread_barrier_wait(&g->bar);
    void * tmp = malloc(n * w + 1);
    for (i = 0; i < n; i++) {
        uint64_t acc = 0;
        for (k = 0; k < g->size; k++)
            acc += (w == 8) ? ((const uint64_t *) g->slot[k])[i] : (uint64_t) ((const uint32_t *) g->slot[k])[i];
        if (w == 8) ((uint64_t *) tmp)[i] = acc; else ((uint32_t *) tmp)[i] = (uint32_t) acc;
    }
    pthread_barrier_wait(&g->bar);      /* everyone has read every input: in-place results may land */
    memcpy(r, tmp, n * w);
    free(tmp);
    return ncclSuccess;
}

ncclResult_t ncclAllGather(const void * s, void * r, size_t n, ncclDataType_t t, ncclComm_t c, cudaStream_t st)
{
    struct mock_group * g = c->g;
    const size_t bytes = n * nccl_size(t);
    int k;
    (void) st;
    if (t_group_depth) return ncclInvalidUsage;
    g->slot[c->rank] = s;
    pthread_barrier_wait(&g->bar);
    /* in place (s == r + rank * bytes) is allowed: a peer's piece lives where I never write */
    for (k = 0; k < g->size; k++)
        if ((const char *) g->slot[k] != (char *) r + (size_t) k * bytes) memmove((char *) r + (size_t) k * bytes, g->slot[k], bytes);
    pthread_barrier_wait(&g->bar);
    return ncclSuccess;
}

ncclResult_t ncclGroupStart(void) { if (t_group_depth++ == 0) { t_nops = 0; t_group_comm = NULL; } return ncclSuccess; }

static ncclResult_t group_add(int kind, const void * s, void * r, size_t bytes, int peer, ncclComm_t c)
{
    if (!t_group_depth) return ncclInvalidUsage;        /* the host code only uses these inside a group */
    if (t_group_comm && t_group_comm != c) return ncclInvalidUsage;
    if (t_nops == MOCK_MAXOPS || peer < 0 || peer >= c->g->size) return ncclInvalidArgument;
    t_group_comm = c;
    t_ops[t_nops].kind = kind; t_ops[t_nops].send = s; t_ops[t_nops].recv = r; t_ops[t_nops].bytes = bytes; t_ops[t_nops].peer = peer;
    t_nops++;
    return ncclSuccess;
}

ncclResult_t ncclBroadcast(const void * s, void * r, size_t n, ncclDataType_t t, int root, ncclComm_t c, cudaStream_t st)
{ (void) st; return group_add(2, s, r, n * nccl_size(t), root, c); }
ncclResult_t ncclSend(const void * s, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t st)
{ (void) st; if (peer == c->rank) return ncclInvalidArgument; return group_add(0, s, NULL, n * nccl_size(t), peer, c); }
ncclResult_t ncclRecv(void * r, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t st)
{ (void) st; if (peer == c->rank) return ncclInvalidArgument; return group_add(1, NULL, r, n * nccl_size(t), peer, c); }

#define FIFO(g, from, to, ch) (&(g)->fifo[(((size_t) (from) * (g)->size + (to)) * 2) + (ch)])

static void fifo_post(struct mock_group * g, int from, int to, int ch, const void * ptr, size_t bytes)
{
    struct mock_fifo * f = FIFO(g, from, to, ch);
    pthread_mutex_lock(&g->mlock);
    while (f->tail - f->head == MOCK_FIFO) pthread_cond_wait(&g->mcond, &g->mlock);
    const unsigned long seq = f->tail++;
    f->msg[seq % MOCK_FIFO].ptr = ptr;
    f->msg[seq % MOCK_FIFO].bytes = bytes;
    pthread_cond_broadcast(&g->mcond);
    pthread_mutex_unlock(&g->mlock);
}

static int fifo_take(struct mock_group * g, int from, int to, int ch, void * dst, size_t bytes)
{
    struct mock_fifo * f = FIFO(g, from, to, ch);
    pthread_mutex_lock(&g->mlock);
    while (f->tail == f->head) pthread_cond_wait(&g->mcond, &g->mlock);
    const void * src = f->msg[f->head % MOCK_FIFO].ptr;
    const size_t have = f->msg[f->head % MOCK_FIFO].bytes;
    pthread_mutex_unlock(&g->mlock);
    const int ok = (have == bytes);
    if (ok && bytes) memcpy(dst, src, bytes);           /* the sender keeps its buffer until head moves */
    else if (!ok) fprintf(stderr, "mock NCCL: rank %d expects %zu bytes from rank %d, which sends %zu\n", to, bytes, from, have);
    pthread_mutex_lock(&g->mlock);
    f->head++;
    pthread_cond_broadcast(&g->mcond);
    pthread_mutex_unlock(&g->mlock);
    return ok;
}

/* only `from` posts to this queue: once it has stopped posting, head == tail means all was read */
static void fifo_wait_drained(struct mock_group * g, int from, int to, int ch)
{
    struct mock_fifo * f = FIFO(g, from, to, ch);
    pthread_mutex_lock(&g->mlock);
    while (f->head != f->tail) pthread_cond_wait(&g->mcond, &g->mlock);
    pthread_mutex_unlock(&g->mlock);
}

ncclResult_t ncclGroupEnd(void)
{
    int i, k, bad = 0;
    if (t_group_depth <= 0) return ncclInvalidUsage;
    if (--t_group_depth > 0) return ncclSuccess;
    struct ncclComm * c = t_group_comm;
    if (!c) return ncclSuccess;                          /* nothing posted */
    struct mock_group * g = c->g;
    const int me = c->rank;
    for (i = 0; i < t_nops; i++) {
        const struct mock_op * o = &t_ops[i];
        if (o->kind == 0) fifo_post(g, me, o->peer, 0, o->send, o->bytes);
        else if (o->kind == 2 && o->peer == me)
            for (k = 0; k < g->size; k++) if (k != me) fifo_post(g, me, k, 1, o->send, o->bytes);
    }
    for (i = 0; i < t_nops; i++) {
        const struct mock_op * o = &t_ops[i];
        if (o->kind == 1) bad |= !fifo_take(g, o->peer, me, 0, o->recv, o->bytes);
        else if (o->kind == 2 && o->peer != me) bad |= !fifo_take(g, o->peer, me, 1, o->recv, o->bytes);
        else if (o->kind == 2 && o->recv != o->send && o->bytes) memmove(o->recv, o->send, o->bytes);
    }
    for (i = 0; i < t_nops; i++) {
        const struct mock_op * o = &t_ops[i];
        if (o->kind == 0) fifo_wait_drained(g, me, o->peer, 0);
        else if (o->kind == 2 && o->peer == me)
            for (k = 0; k < g->size; k++) if (k != me) fifo_wait_drained(g, me, k, 1);
    }
    return bad ? ncclInvalidUsage : ncclSuccess;
}

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
        uint32_t * hist, uint64_t * diff, const void * ref, mpsk_stream_t stream)
{
    size_t i;
    uint32_t q;
    (void) stream;
    if (n == 0) return 0;
    if ((elsize != 8 && elsize != 16) || (nh != 4 && nh != 8) || d0 + nh > 8) return (int) cudaErrorInvalidValue;
    LAUNCHED();
    const uint64_t k0 = rec_key(ref ? ref : recs, 0, elsize, khi, flip);
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

size_t mpsk_fixup_tile_items(void) { return 2048; }

int mpsk_fixup_rec(void * recs, size_t n, size_t elsize, int khi, uint64_t flip, uint32_t lobits,
        uint32_t * worklist, uint32_t * nwork, uint32_t cap, size_t tile0, size_t ntiles, mpsk_stream_t stream)
{
    size_t i = 0;
    (void) stream;
    if (n == 0 || tile0 * 2048 >= n) return 0;
    LAUNCHED();
    /* runs whose first record lies in tiles [tile0, tile0 + ntiles) */
    const size_t first = tile0 * 2048, last = ntiles ? (tile0 + ntiles) * 2048 : n;
    /* back up to the head of the run that `first` sits in: that run belongs to an earlier tile */
    i = first;
    while (i > 0 && i < n && (lobits >= 64 ? 0 : rec_key(recs, i - 1, elsize, khi, flip) >> lobits)
                             == (lobits >= 64 ? 0 : rec_key(recs, i, elsize, khi, flip) >> lobits)) i--;
    if (i < first) {                       /* skip it */
        const uint64_t hi0 = lobits >= 64 ? 0 : rec_key(recs, i, elsize, khi, flip) >> lobits;
        while (i < n && (lobits >= 64 ? 0 : rec_key(recs, i, elsize, khi, flip) >> lobits) == hi0) i++;
    }
    const uint64_t lomask = lobits >= 64 ? ~0ULL : ((1ULL << lobits) - 1ULL);
    char * tmp = (char *) malloc(256 * elsize);
    while (i < n && i < last) {
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

static uint64_t mock_mix64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static int cmp_u64(const void * a, const void * b)
{
    const uint64_t x = *(const uint64_t *) a, y = *(const uint64_t *) b;
    return (x > y) - (x < y);
}

int mpsk_prefix_pairs(const void * recs, size_t n, size_t elsize, uint32_t s, int khi, uint64_t flip,
        const uint32_t * lobits, uint32_t nl, uint64_t * table, uint32_t log2_tsize, uint64_t * pairs, mpsk_stream_t stream)
{
    /* same contract, by sorting the sampled high parts (the table is left alone) */
    uint32_t i, j;
    (void) stream; (void) table; (void) log2_tsize;
    if (s == 0 || nl == 0 || n == 0) return 0;
    if (nl > 2) return 1;
    LAUNCHED();
    uint64_t * v = (uint64_t *) malloc(sizeof(uint64_t) * s);
    for (j = 0; j <= nl; j++) {
        for (i = 0; i < s; i++) {
            const size_t pos = (size_t) (mock_mix64(0x5EED5A3Bu + i) % (uint64_t) n);
            uint64_t k;
            memcpy(&k, (const unsigned char *) recs + pos * elsize + ((khi && elsize == 16) ? 8 : 0), 8);
            k ^= flip;
            v[i] = j == nl ? (uint64_t) pos : (lobits[j] >= 64 ? 0 : k >> lobits[j]);      /* [nl]: positions drawn twice */
        }
        qsort(v, s, sizeof(uint64_t), cmp_u64);
        uint64_t c = 0;
        uint32_t run = 0;
        for (i = 1; i < s; i++) {
            run = (v[i] == v[i - 1]) ? run + 1 : 0;
            c += run;
        }
        pairs[j] += c;
    }
    free(v);
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

/* the one-kernel descent over peer mailboxes (candidate): the PROTOCOL of splitter_descent_peer_kernel restated with
 * rank threads -- per level, counts into the own mailbox (parity = level & 1), release-store of the per-splitter flag,
 * acquire-poll of the peers' flags, sum of the peers' counts, digit choice. Layout as in kernels/splitter.cuh:
 * counts[2][63][256] u64, then 256 flag words. What it cannot show: the GPU memory model, co-residency. */
#define MOCK_PEER_MAXS 63
size_t mpsk_peer_box_bytes(void) { return (size_t) 2 * MOCK_PEER_MAXS * 256 * sizeof(uint64_t) + 256 * sizeof(uint32_t); }
int mpsk_splitter_descent_peer(struct mpsk_keyview view, size_t n, uint32_t nw, uint64_t * prefix, const uint64_t * target,
        int nsplit, int level0, int nlevels, uint32_t me, uint32_t p, void * const * boxes, uint32_t seq, uint32_t * err,
        mpsk_stream_t stream)
{
    int level, b;
    uint32_t d, w, r;
    (void) stream;
    if (nsplit <= 0 || level0 >= nlevels) return 0;
    if (nw > 16 || nsplit > MOCK_PEER_MAXS || p > 64 || me >= p) return (int) cudaErrorInvalidValue;
    LAUNCHED();
    if (getenv("MOCK_TRACE")) fprintf(stderr, "mock: peer descent, rank %u of %u, %d splitters, levels %d..%d, seq %u\n", me, p, nsplit, level0, nlevels - 1, seq);
    uint64_t * mycounts = (uint64_t *) boxes[me];
    uint32_t * myflags = (uint32_t *) (mycounts + (size_t) 2 * MOCK_PEER_MAXS * 256);
    for (level = level0; level < nlevels; level++) {
        const uint32_t par = (uint32_t) level & 1u, want = seq + (uint32_t) level + 1u;
        const uint32_t byteidx = 8 * nw - 1 - (uint32_t) level, wi = byteidx >> 3, sh = (byteidx & 7) * 8;
        for (b = 0; b < nsplit; b++) {
            for (d = 0; d < 256; d++) {
                uint64_t cand[16];
                for (w = 0; w < nw; w++) {
                    uint64_t x = prefix[(size_t) b * nw + w];
                    if (w < wi) x = ~0ULL;
                    else if (w == wi) x |= ((uint64_t) d << sh) | ((sh == 0) ? 0ULL : ((1ULL << sh) - 1ULL));
                    cand[w] = x;
                }
                mycounts[((size_t) par * MOCK_PEER_MAXS + b) * 256 + d] = bound_view(view, n, cand, nw, 1);
            }
            __atomic_store_n(&myflags[b], want, __ATOMIC_RELEASE);
        }
        for (b = 0; b < nsplit; b++) {
            uint32_t pick = 255;
            for (r = 0; r < p; r++) {
                if (r == me) continue;
                const uint32_t * pf = (const uint32_t *) ((const uint64_t *) boxes[r] + (size_t) 2 * MOCK_PEER_MAXS * 256) + b;
                const double t0 = now_ms();
                while ((int32_t) (__atomic_load_n(pf, __ATOMIC_ACQUIRE) - want) < 0)
                    if (now_ms() - t0 > 20000.0) { *err = 1; return 0; }
            }
            for (d = 0; d < 256; d++) {
                uint64_t sum = 0;
                for (r = 0; r < p; r++) sum += ((const uint64_t *) boxes[r])[((size_t) par * MOCK_PEER_MAXS + b) * 256 + d];
                if (sum >= target[b]) { pick = d; break; }
            }
            prefix[(size_t) b * nw + wi] |= ((uint64_t) pick) << sh;
        }
    }
    return 0;
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

/* record i (an index into the receive buffer's layout) of run r: the one run that was not copied is read from
 * its own base with the same index (mpsort_kernels.h) */
static const unsigned char * merge_rec(const void * recv, uint32_t r, uint32_t self_run, const void * self_recv, size_t i, size_t elsize)
{
    const uintptr_t base = (self_recv && r == self_run) ? (uintptr_t) self_recv : (uintptr_t) recv;
    return (const unsigned char *) (base + i * elsize);
}

int mpsk_merge_samples(const void * recv, size_t elsize, size_t offset, uint32_t width, uint32_t nwords, int is_signed,
        uint32_t p, uint32_t S, uint32_t k, const uint32_t * rdispl, const uint32_t * sstart, uint32_t self_run, const void * self_recv,
        uint64_t * skeys, mpsk_stream_t stream)
{
    uint32_t r, j;
    (void) k; (void) stream;
    if (sstart[p] == 0) return 0;
    LAUNCHED();
    for (r = 0; r < p; r++)
        for (j = 0; j < sstart[r + 1] - sstart[r]; j++) {
            const size_t pos = (size_t) rdispl[r] + (size_t) (j + 1) * S - 1;
            skeys[sstart[r] + j] = pack_word(merge_rec(recv, r, self_run, self_recv, pos, elsize), offset, width, nwords, is_signed, 0);
        }
    return 0;
}

/* the samples in (key, run, position) order: same contract as the kernel, as a p-way merge of the runs' sorted lists */
int mpsk_merge_rank_samples(const uint64_t * skeys, uint32_t p, const uint32_t * sstart, uint64_t * sorted_skeys,
        uint32_t * sorted_sid, mpsk_stream_t stream)
{
    uint32_t head[64], r, o;
    (void) stream;
    if (p > 32) return 1;
    if (sstart[p] == 0) return 0;
    LAUNCHED();
    for (r = 0; r < p; r++) head[r] = sstart[r];
    for (o = 0; o < sstart[p]; o++) {
        int best = -1;
        for (r = 0; r < p; r++)
            if (head[r] < sstart[r + 1] && (best < 0 || skeys[head[r]] < skeys[head[best]])) best = (int) r;
        sorted_skeys[o] = skeys[head[best]];
        sorted_sid[o] = head[best];
        head[best]++;
    }
    return 0;
}

int mpsk_merge_runs(const void * recv, void * out, size_t elsize, size_t offset, uint32_t width, uint32_t nwords, int is_signed,
        uint32_t p, uint32_t S, uint32_t k, const uint32_t * rdispl, const uint32_t * sstart, uint32_t self_run, const void * self_recv,
        const uint64_t * sorted_skeys, const uint32_t * sorted_sid, uint32_t ntiles, uint32_t * cut, uint32_t * overflow,
        mpsk_stream_t stream)
{
    uint32_t head[64], r;
    size_t o = 0;
    (void) S; (void) k; (void) sstart; (void) sorted_skeys; (void) sorted_sid; (void) ntiles; (void) cut; (void) overflow; (void) stream;
    if (p > 32 || (self_recv && (self_run >= p || rdispl[p] >= 0x80000000u))) return (int) cudaErrorInvalidValue;
    LAUNCHED();
    for (r = 0; r < p; r++) head[r] = rdispl[r];
    for (;;) {
        int best = -1;
        uint64_t kb = 0;
        for (r = 0; r < p; r++) {
            if (head[r] >= rdispl[r + 1]) continue;
            const uint64_t kr = pack_word(merge_rec(recv, r, self_run, self_recv, head[r], elsize), offset, width, nwords, is_signed, 0);
            if (best < 0 || kr < kb) { best = (int) r; kb = kr; }      /* ties: the lower run */
        }
        if (best < 0) break;
        memcpy((char *) out + o * elsize, merge_rec(recv, (uint32_t) best, self_run, self_recv, head[best], elsize), elsize);
        head[best]++;
        o++;
    }
    return 0;
}

int mpsk_p2p_alltoallv(const void * const * src, void * const * dst, const uint64_t * bytes, const unsigned char * remote,
        int nseg, mpsk_stream_t stream)
{
    int k;
    (void) remote; (void) stream;
    LAUNCHED();
    for (k = 0; k < nseg; k++) if (bytes[k]) memcpy(dst[k], src[k], (size_t) bytes[k]);
    return 0;
}

int mpsk_p2p_gather_alltoallv(const void * base, const uint32_t * const * idx, void * const * dst, const uint64_t * nrec,
        size_t elsize, int nseg, mpsk_stream_t stream)
{
    int k;
    uint64_t i;
    (void) stream;
    LAUNCHED();
    for (k = 0; k < nseg; k++)
        for (i = 0; i < nrec[k]; i++)
            memcpy((char *) dst[k] + (size_t) i * elsize, (const char *) base + (size_t) idx[k][i] * elsize, elsize);
    return 0;
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

int mpsk_multiset_hash(const void * base, size_t n, size_t elsize, uint64_t * out, mpsk_stream_t stream)
{
    size_t i, b, k;
    (void) stream;
    if (n == 0 || elsize == 0) return 0;
    LAUNCHED();
    for (i = 0; i < n; i++) {
        const unsigned char * r = (const unsigned char *) base + i * elsize;
        uint64_t h = 0x243F6A8885A308D3ULL;
        for (b = 0; b < elsize; b += 8) {
            uint64_t w = 0;
            for (k = 0; k < 8 && b + k < elsize; k++) w |= (uint64_t) r[b + k] << (8 * k);
            h = mock_mix64(h ^ w);
        }
        out[0] += h;
        out[1] ^= h;
    }
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
