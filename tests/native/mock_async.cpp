/*
 * mock_async.cpp -- TEST INFRASTRUCTURE ONLY: stream semantics for the mock device.
 *
 * mock_device.c runs every stream-ordered call at the moment it is made, which hides exactly the
 * bugs a multi-stream host flow can have: a consumer that does not wait for its producer. This layer
 * exports the stream-ordered entry points (CUDA copies / memsets / events, the NCCL calls that take a
 * stream, every launch of the kernel ABI) and decides WHEN each one runs:
 *
 *   - every call becomes an operation in the FIFO of its stream; nothing runs when it is enqueued;
 *   - an operation runs only when something legally forces it: cudaStreamSynchronize /
 *     cudaEventSynchronize on the host, a cudaStreamWaitEvent of another stream whose event is not
 *     complete yet, a blocking copy, cudaFree, stream destruction;
 *   - besides that a seeded scheduler (MOCK_ASYNC=<seed>) runs random runnable operations of the
 *     calling thread's streams at random moments, so that different seeds walk through different
 *     LEGAL interleavings of the streams.
 * A missing dependency (a merge that does not wait for the transfer it reads, a staging buffer reused
 * before its copy ran, a host read before the synchronisation) then shows up as wrong bytes in the
 * tests' comparison with the oracle, on the CPU box. Without MOCK_ASYNC every operation runs at once,
 * as in the plain mock.
 *
 * Modelled after the CUDA rules the host code relies on: operations of one stream run in order;
 * cudaStreamWaitEvent waits for the event's most recent record AT THE TIME OF THE CALL; copies to or
 * from pageable host memory return when the host buffer may be reused / holds the data, copies with
 * pinned memory (cudaMallocHost) are fully asynchronous; arguments of a launch are captured at the
 * launch (host arrays of pointers and counts included), device memory is read when the kernel runs;
 * cudaFree waits for the device.  NCCL operations block, when they RUN, until the other ranks run
 * theirs (threads of this process, see mock_device.c).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <deque>
#include <functional>
#include <memory>
#include <map>
#include <vector>

#include <cuda_runtime_api.h>
#include <nccl.h>

#include "mpsort_kernels.h"

extern "C" {
/* the synchronous implementations of mock_device.c (built with -DMOCK_WITH_ASYNC_LAYER) */
cudaError_t mocksync_cudaMemcpyAsync(void *, const void *, size_t, enum cudaMemcpyKind, cudaStream_t);
cudaError_t mocksync_cudaMemsetAsync(void *, int, size_t, cudaStream_t);
cudaError_t mocksync_cudaFree(void *);
ncclResult_t mocksync_ncclAllReduce(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
ncclResult_t mocksync_ncclAllGather(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
ncclResult_t mocksync_ncclBroadcast(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
ncclResult_t mocksync_ncclSend(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
ncclResult_t mocksync_ncclRecv(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
ncclResult_t mocksync_ncclGroupStart(void);
ncclResult_t mocksync_ncclGroupEnd(void);
#define K(name) mocksync_##name
int K(mpsk_extract_keys)(const void *, size_t, size_t, size_t, uint32_t, uint32_t, int, uint32_t, uint64_t, uint64_t *, uint32_t *, uint64_t *, mpsk_stream_t);
int K(mpsk_rec_histograms)(const void *, size_t, size_t, int, uint64_t, uint32_t, uint32_t, uint32_t *, uint64_t *, const void *, mpsk_stream_t);
int K(mpsk_rec_sample_diff)(const void *, size_t, size_t, int, uint32_t, uint64_t *, mpsk_stream_t);
int K(mpsk_scan_histograms)(const uint32_t *, uint32_t *, int, mpsk_stream_t);
int K(mpsk_onesweep_pass)(const uint64_t *, const uint32_t *, uint64_t *, uint32_t *, size_t, int, const uint32_t *, void *, mpsk_stream_t);
int K(mpsk_onesweep_pass_rec)(const void *, void *, size_t, size_t, int, int, uint64_t, const uint32_t *, void *, mpsk_stream_t);
int K(mpsk_fixup_rec)(void *, size_t, size_t, int, uint64_t, uint32_t, uint32_t *, uint32_t *, uint32_t, size_t, size_t, mpsk_stream_t);
int K(mpsk_fixup_extents)(const void *, size_t, size_t, int, uint64_t, uint32_t, const uint32_t *, uint32_t, uint32_t *, mpsk_stream_t);
int K(mpsk_prefix_pairs)(const void *, size_t, size_t, uint32_t, int, uint64_t, const uint32_t *, uint32_t, uint64_t *, uint32_t, uint64_t *, mpsk_stream_t);
int K(mpsk_gather_u64)(const uint64_t *, const uint32_t *, uint64_t *, size_t, mpsk_stream_t);
int K(mpsk_gather_records)(const void *, const uint32_t *, void *, size_t, size_t, mpsk_stream_t);
int K(mpsk_splitter_count)(struct mpsk_keyview, size_t, uint32_t, const uint64_t *, int, int, uint64_t *, mpsk_stream_t);
int K(mpsk_splitter_select)(const uint64_t *, const uint64_t *, uint64_t *, uint32_t, int, int, mpsk_stream_t);
int K(mpsk_splitter_final)(struct mpsk_keyview, size_t, uint32_t, const uint64_t *, int, uint64_t *, mpsk_stream_t);
int K(mpsk_splitter_descent_peer)(struct mpsk_keyview, size_t, uint32_t, uint64_t *, const uint64_t *, int, int, int, uint32_t, uint32_t, void * const *, uint32_t, uint32_t *, mpsk_stream_t);
int K(mpsk_sum_u64)(uint64_t *, const uint64_t * const *, int, size_t, mpsk_stream_t);
int K(mpsk_merge_samples)(const void *, size_t, size_t, uint32_t, uint32_t, int, uint32_t, uint32_t, uint32_t, const uint32_t *, const uint32_t *, uint32_t, const void *, uint64_t *, mpsk_stream_t);
int K(mpsk_merge_rank_samples)(const uint64_t *, uint32_t, const uint32_t *, uint64_t *, uint32_t *, mpsk_stream_t);
int K(mpsk_merge_runs)(const void *, void *, size_t, size_t, uint32_t, uint32_t, int, uint32_t, uint32_t, uint32_t, const uint32_t *, const uint32_t *, uint32_t, const void *, const uint64_t *, const uint32_t *, uint32_t, uint32_t *, uint32_t *, mpsk_stream_t);
int K(mpsk_p2p_alltoallv)(const void * const *, void * const *, const uint64_t *, const unsigned char *, int, mpsk_stream_t);
int K(mpsk_p2p_gather_alltoallv)(const void *, const uint32_t * const *, void * const *, const uint64_t *, size_t, int, mpsk_stream_t);
int K(mpsk_checksum)(const void *, size_t, uint64_t *, mpsk_stream_t);
int K(mpsk_multiset_hash)(const void *, size_t, size_t, uint64_t *, mpsk_stream_t);
int K(mpsk_generate)(void *, size_t, size_t, int, uint64_t, uint64_t, uint64_t, mpsk_stream_t);
int K(mpsk_check_sorted)(const void *, size_t, size_t, size_t, uint32_t, uint32_t, int, int, size_t, uint64_t *, uint64_t *, mpsk_stream_t);
}

/* ------------------------------------------------------------------------- */
/* streams, events, the scheduler                                             */

struct MockEvent;
struct Op {
    std::function<void()> run;
    MockEvent * wait_ev = nullptr;      /* a cudaStreamWaitEvent: runnable once wait_ev->done_gen >= wait_gen */
    uint64_t wait_gen = 0;
};
struct MockStream { std::deque<Op> q; };
struct MockEvent {
    MockStream * stream = nullptr;      /* where the most recent record was enqueued */
    uint64_t gen = 0, done_gen = 0;     /* records enqueued / records that have run */
    double t_ms = 0.0;
};

static unsigned long g_seed = 0;
static pthread_mutex_t g_reg_lock = PTHREAD_MUTEX_INITIALIZER;
static std::map<const void *, size_t> g_pinned;     /* cudaMallocHost allocations: base -> bytes */
static unsigned g_thread_counter = 0;

static thread_local std::vector<MockStream *> t_streams;
static thread_local uint64_t t_rng = 0;
static thread_local int t_depth = 0;

static bool read_switch()
{
    const char * e = getenv("MOCK_ASYNC");
    g_seed = e ? strtoul(e, NULL, 10) : 0;
    return e && *e;
}
static const bool g_async = read_switch();      /* read once, when the library is loaded */
static bool async_on() { return g_async; }

static uint64_t rnd()
{
    if (t_rng == 0) {
        pthread_mutex_lock(&g_reg_lock);
        t_rng = 0x9E3779B97F4A7C15ULL * (g_seed + 1) + 0xD1B54A32D192ED03ULL * (++g_thread_counter);
        pthread_mutex_unlock(&g_reg_lock);
    }
    t_rng ^= t_rng << 13; t_rng ^= t_rng >> 7; t_rng ^= t_rng << 17;
    return t_rng;
}

static double now_ms()
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return 1e3 * (double) t.tv_sec + 1e-6 * (double) t.tv_nsec;
}

static void run_one(MockStream * s);

/* run operations of stream s until `done()` holds or the stream is empty */
static void drive(MockStream * s, const std::function<bool()> & done)
{
    if (++t_depth > 64) { fprintf(stderr, "mock: streams wait for each other in a cycle\n"); abort(); }
    while (!done() && !s->q.empty()) run_one(s);
    --t_depth;
}

static void run_one(MockStream * s)
{
    Op op = std::move(s->q.front());
    if (op.wait_ev && op.wait_ev->done_gen < op.wait_gen) {
        /* the awaited record has not run: its stream must get there first. The wait stays at the head of s. */
        MockEvent * e = op.wait_ev;
        const uint64_t gen = op.wait_gen;
        s->q.front() = std::move(op);
        if (!e->stream || e->stream == s) { fprintf(stderr, "mock: a stream waits for an event it has not reached itself\n"); abort(); }
        drive(e->stream, [e, gen] { return e->done_gen >= gen; });
        if (e->done_gen < gen) { fprintf(stderr, "mock: awaited event can never complete\n"); abort(); }
        s->q.pop_front();
        return;
    }
    s->q.pop_front();
    if (op.run) op.run();
}

static bool head_runnable(MockStream * s)
{
    if (s->q.empty()) return false;
    const Op & op = s->q.front();
    return !(op.wait_ev && op.wait_ev->done_gen < op.wait_gen);
}

/* the seeded scheduler: now and then, run a few runnable operations of this thread's streams */
static void random_progress()
{
    if (!async_on() || t_streams.empty()) return;
    if (rnd() % 3) return;
    int steps = (int) (rnd() % 6);
    while (steps-- > 0) {
        MockStream * s = t_streams[rnd() % t_streams.size()];
        if (head_runnable(s)) run_one(s);
    }
}

static void enqueue(cudaStream_t st, std::function<void()> f)
{
    MockStream * s = (MockStream *) st;
    if (!async_on() || !s) { f(); return; }
    Op op;
    op.run = std::move(f);
    s->q.push_back(std::move(op));
    random_progress();
}

static void sync_stream(MockStream * s) { if (s) drive(s, [] { return false; }); }
static void sync_thread() { for (MockStream * s : t_streams) sync_stream(s); }

static void must(int rc, const char * what)
{
    if (rc != 0) { fprintf(stderr, "mock: deferred %s failed with %d\n", what, rc); abort(); }
}

template <typename T> static std::shared_ptr<std::vector<T>> snap(const T * p, size_t n)
{
    return std::shared_ptr<std::vector<T>>(new std::vector<T>(p, p + n));
}

static bool is_pinned(const void * p)
{
    bool yes = false;
    pthread_mutex_lock(&g_reg_lock);
    auto it = g_pinned.upper_bound(p);
    if (it != g_pinned.begin()) {
        --it;
        yes = (const char *) p < (const char *) it->first + it->second;
    }
    pthread_mutex_unlock(&g_reg_lock);
    return yes;
}

extern "C" {

cudaError_t cudaStreamCreateWithFlags(cudaStream_t * s, unsigned f)
{
    (void) f;
    MockStream * m = new MockStream();
    t_streams.push_back(m);
    *s = (cudaStream_t) m;
    return cudaSuccess;
}

cudaError_t cudaStreamDestroy(cudaStream_t s)
{
    MockStream * m = (MockStream *) s;
    if (!m) return cudaSuccess;
    sync_stream(m);
    for (size_t i = 0; i < t_streams.size(); i++) if (t_streams[i] == m) { t_streams.erase(t_streams.begin() + i); break; }
    delete m;
    return cudaSuccess;
}

cudaError_t cudaStreamSynchronize(cudaStream_t s) { sync_stream((MockStream *) s); return cudaSuccess; }

cudaError_t cudaEventCreate(cudaEvent_t * e) { *e = (cudaEvent_t) new MockEvent(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t * e, unsigned f) { (void) f; return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e)
{
    MockEvent * m = (MockEvent *) e;
    if (m && m->stream && m->done_gen < m->gen) { MockEvent * ev = m; drive(m->stream, [ev] { return ev->done_gen >= ev->gen; }); }
    delete m;
    return cudaSuccess;
}

cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s)
{
    MockEvent * m = (MockEvent *) e;
    const uint64_t gen = ++m->gen;
    m->stream = (MockStream *) s;
    enqueue(s, [m, gen] { if (m->done_gen < gen) m->done_gen = gen; m->t_ms = now_ms(); });
    return cudaSuccess;
}

cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned f)
{
    MockEvent * m = (MockEvent *) e;
    MockStream * st = (MockStream *) s;
    (void) f;
    if (!async_on() || !st || m->gen == 0 || m->done_gen >= m->gen) return cudaSuccess;
    Op op;
    op.wait_ev = m;
    op.wait_gen = m->gen;            /* the most recent record at the time of THIS call */
    st->q.push_back(std::move(op));
    random_progress();
    return cudaSuccess;
}

cudaError_t cudaEventSynchronize(cudaEvent_t e)
{
    MockEvent * m = (MockEvent *) e;
    if (m->stream && m->done_gen < m->gen) drive(m->stream, [m] { return m->done_gen >= m->gen; });
    return cudaSuccess;
}

cudaError_t cudaEventElapsedTime(float * ms, cudaEvent_t a, cudaEvent_t b)
{
    MockEvent * x = (MockEvent *) a, * y = (MockEvent *) b;
    if (x->gen == 0 || y->gen == 0) return cudaErrorInvalidResourceHandle;
    if (x->done_gen < x->gen || y->done_gen < y->gen) return cudaErrorNotReady;
    const double d = y->t_ms - x->t_ms;
    *ms = (float) (d > 1e-3 ? d : 1e-3);
    return cudaSuccess;
}

cudaError_t cudaMallocHost(void ** p, size_t n)
{
    *p = malloc(n ? n : 1);
    if (!*p) return cudaErrorMemoryAllocation;
    pthread_mutex_lock(&g_reg_lock);
    g_pinned[*p] = n ? n : 1;
    pthread_mutex_unlock(&g_reg_lock);
    return cudaSuccess;
}

cudaError_t cudaFreeHost(void * p)
{
    sync_thread();
    pthread_mutex_lock(&g_reg_lock);
    g_pinned.erase(p);
    pthread_mutex_unlock(&g_reg_lock);
    free(p);
    return cudaSuccess;
}

/* cudaFree waits for the device: everything this thread has queued runs first */
cudaError_t cudaFree(void * p) { sync_thread(); return mocksync_cudaFree(p); }

/* blocking copies / memsets: ordered after nothing but themselves (the host code's streams are non-blocking) */
cudaError_t cudaMemcpy(void * d, const void * s, size_t n, enum cudaMemcpyKind k) { (void) k; if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void * d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }

cudaError_t cudaMemcpyAsync(void * d, const void * s, size_t n, enum cudaMemcpyKind k, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    if (!async_on() || !st) return mocksync_cudaMemcpyAsync(d, s, n, k, st);
    struct cudaPointerAttributes as, ad;
    cudaPointerGetAttributes(&as, s);
    cudaPointerGetAttributes(&ad, d);
    const bool src_dev = as.type == cudaMemoryTypeDevice, dst_dev = ad.type == cudaMemoryTypeDevice;
    if (!src_dev && dst_dev && !is_pinned(s)) {
        /* pageable source: staged before the call returns, the caller may reuse it at once */
        std::shared_ptr<std::vector<char>> snap(new std::vector<char>((const char *) s, (const char *) s + n));
        enqueue(st, [d, snap, n] { memcpy(d, snap->data(), n); });
        return cudaSuccess;
    }
    enqueue(st, [d, s, n] { memmove(d, s, n); });
    if (src_dev && !dst_dev && !is_pinned(d)) sync_stream((MockStream *) st);     /* pageable destination: returns with the data */
    return cudaSuccess;
}

cudaError_t cudaMemsetAsync(void * d, int v, size_t n, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    enqueue(st, [d, v, n] { memset(d, v, n); });
    return cudaSuccess;
}

/* ------------------------------------------------------------------------- */
/* NCCL calls that take a stream                                              */

struct GroupOp { int kind; const void * s; void * r; size_t n; ncclDataType_t t; int peer; ncclComm_t c; };
static thread_local int t_group = 0;
static thread_local std::vector<GroupOp> t_gops;
static thread_local cudaStream_t t_gstream = nullptr;

ncclResult_t ncclAllReduce(const void * s, void * r, size_t n, ncclDataType_t t, ncclRedOp_t o, ncclComm_t c, cudaStream_t st)
{
    if (t_group) return ncclInvalidUsage;
    enqueue(st, [=] { must((int) mocksync_ncclAllReduce(s, r, n, t, o, c, st), "ncclAllReduce"); });
    return ncclSuccess;
}

ncclResult_t ncclAllGather(const void * s, void * r, size_t n, ncclDataType_t t, ncclComm_t c, cudaStream_t st)
{
    if (t_group) return ncclInvalidUsage;
    enqueue(st, [=] { must((int) mocksync_ncclAllGather(s, r, n, t, c, st), "ncclAllGather"); });
    return ncclSuccess;
}

ncclResult_t ncclGroupStart(void) { if (t_group++ == 0) { t_gops.clear(); t_gstream = nullptr; } return ncclSuccess; }

static ncclResult_t group_add(int kind, const void * s, void * r, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t st)
{
    if (!t_group) return ncclInvalidUsage;
    if (!t_gops.empty() && t_gstream != st) { fprintf(stderr, "mock: one NCCL group over two streams is not modelled\n"); abort(); }
    t_gstream = st;
    t_gops.push_back(GroupOp{ kind, s, r, n, t, peer, c });
    return ncclSuccess;
}
ncclResult_t ncclBroadcast(const void * s, void * r, size_t n, ncclDataType_t t, int root, ncclComm_t c, cudaStream_t st) { return group_add(2, s, r, n, t, root, c, st); }
ncclResult_t ncclSend(const void * s, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t st) { return group_add(0, s, NULL, n, t, peer, c, st); }
ncclResult_t ncclRecv(void * r, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t st) { return group_add(1, NULL, r, n, t, peer, c, st); }

ncclResult_t ncclGroupEnd(void)
{
    if (t_group <= 0) return ncclInvalidUsage;
    if (--t_group > 0) return ncclSuccess;
    if (t_gops.empty()) return ncclSuccess;
    std::vector<GroupOp> ops = t_gops;
    cudaStream_t st = t_gstream;
    enqueue(st, [ops, st] {
        must((int) mocksync_ncclGroupStart(), "ncclGroupStart");
        for (const GroupOp & o : ops) {
            if (o.kind == 0) must((int) mocksync_ncclSend(o.s, o.n, o.t, o.peer, o.c, st), "ncclSend");
            else if (o.kind == 1) must((int) mocksync_ncclRecv(o.r, o.n, o.t, o.peer, o.c, st), "ncclRecv");
            else must((int) mocksync_ncclBroadcast(o.s, o.r, o.n, o.t, o.peer, o.c, st), "ncclBroadcast");
        }
        must((int) mocksync_ncclGroupEnd(), "ncclGroupEnd");
    });
    return ncclSuccess;
}

/* ------------------------------------------------------------------------- */
/* the kernel ABI: a launch captures its arguments, the body runs when the stream gets there */

#define LAUNCH(st, call) do { cudaStream_t st__ = (cudaStream_t) (st); enqueue(st__, [=] { must(call, #call); }); return 0; } while (0)

int mpsk_extract_keys(const void * base, size_t n, size_t elsize, size_t offset, uint32_t width, uint32_t nwords, int is_signed,
        uint32_t g, uint64_t sub, uint64_t * kout, uint32_t * hist, uint64_t * minmax, mpsk_stream_t stream)
{ LAUNCH(stream, K(mpsk_extract_keys)(base, n, elsize, offset, width, nwords, is_signed, g, sub, kout, hist, minmax, stream)); }

int mpsk_rec_histograms(const void * recs, size_t n, size_t elsize, int khi, uint64_t flip, uint32_t d0, uint32_t nh,
        uint32_t * hist, uint64_t * diff, const void * ref, mpsk_stream_t stream)
{
    if (n && ((elsize != 8 && elsize != 16) || (nh != 4 && nh != 8) || d0 + nh > 8)) return (int) cudaErrorInvalidValue;
    LAUNCH(stream, K(mpsk_rec_histograms)(recs, n, elsize, khi, flip, d0, nh, hist, diff, ref, stream));
}

int mpsk_rec_sample_diff(const void * recs, size_t n, size_t elsize, int khi, uint32_t s, uint64_t * diff, mpsk_stream_t stream)
{ LAUNCH(stream, K(mpsk_rec_sample_diff)(recs, n, elsize, khi, s, diff, stream)); }

int mpsk_scan_histograms(const uint32_t * hist, uint32_t * bins, int nhist, mpsk_stream_t stream)
{ LAUNCH(stream, K(mpsk_scan_histograms)(hist, bins, nhist, stream)); }

int mpsk_onesweep_pass(const uint64_t * kin, const uint32_t * vin, uint64_t * kout, uint32_t * vout, size_t n, int shift,
        const uint32_t * bins, void * scratch, mpsk_stream_t stream)
{
    if (n > MPSK_MAX_ITEMS) return (int) cudaErrorInvalidValue;
    LAUNCH(stream, K(mpsk_onesweep_pass)(kin, vin, kout, vout, n, shift, bins, scratch, stream));
}

int mpsk_onesweep_pass_rec(const void * in, void * out, size_t n, size_t elsize, int shift, int khi, uint64_t flip,
        const uint32_t * bins, void * scratch, mpsk_stream_t stream)
{
    if (n && elsize != 8 && elsize != 16) return (int) cudaErrorInvalidValue;
    LAUNCH(stream, K(mpsk_onesweep_pass_rec)(in, out, n, elsize, shift, khi, flip, bins, scratch, stream));
}

int mpsk_fixup_rec(void * recs, size_t n, size_t elsize, int khi, uint64_t flip, uint32_t lobits, uint32_t * worklist,
        uint32_t * nwork, uint32_t cap, size_t tile0, size_t ntiles, mpsk_stream_t stream)
{ LAUNCH(stream, K(mpsk_fixup_rec)(recs, n, elsize, khi, flip, lobits, worklist, nwork, cap, tile0, ntiles, stream)); }

int mpsk_fixup_extents(const void * recs, size_t n, size_t elsize, int khi, uint64_t flip, uint32_t lobits,
        const uint32_t * worklist, uint32_t nwork, uint32_t * lengths, mpsk_stream_t stream)
{ LAUNCH(stream, K(mpsk_fixup_extents)(recs, n, elsize, khi, flip, lobits, worklist, nwork, lengths, stream)); }

int mpsk_prefix_pairs(const void * recs, size_t n, size_t elsize, uint32_t s, int khi, uint64_t flip,
        const uint32_t * lobits, uint32_t nl, uint64_t * table, uint32_t log2_tsize, uint64_t * pairs, mpsk_stream_t stream)
{
    if (nl > 2) return (int) cudaErrorInvalidValue;
    auto lb = snap(lobits, (size_t) (nl ? nl : 1));
    LAUNCH(stream, K(mpsk_prefix_pairs)(recs, n, elsize, s, khi, flip, lb->data(), nl, table, log2_tsize, pairs, stream));
}

int mpsk_gather_u64(const uint64_t * src, const uint32_t * idx, uint64_t * dst, size_t n, mpsk_stream_t stream)
{ LAUNCH(stream, K(mpsk_gather_u64)(src, idx, dst, n, stream)); }

int mpsk_gather_records(const void * base, const uint32_t * idx, void * out, size_t n, size_t elsize, mpsk_stream_t stream)
{ LAUNCH(stream, K(mpsk_gather_records)(base, idx, out, n, elsize, stream)); }

int mpsk_splitter_count(struct mpsk_keyview view, size_t n, uint32_t nw, const uint64_t * prefix, int nsplit, int level,
        uint64_t * counts, mpsk_stream_t stream)
{
    if (nsplit > 0 && nw > 16) return (int) cudaErrorInvalidValue;
    LAUNCH(stream, K(mpsk_splitter_count)(view, n, nw, prefix, nsplit, level, counts, stream));
}

int mpsk_splitter_select(const uint64_t * counts, const uint64_t * target, uint64_t * prefix, uint32_t nw, int nsplit, int level,
        mpsk_stream_t stream)
{ LAUNCH(stream, K(mpsk_splitter_select)(counts, target, prefix, nw, nsplit, level, stream)); }

int mpsk_splitter_final(struct mpsk_keyview view, size_t n, uint32_t nw, const uint64_t * prefix, int nsplit, uint64_t * out,
        mpsk_stream_t stream)
{ LAUNCH(stream, K(mpsk_splitter_final)(view, n, nw, prefix, nsplit, out, stream)); }

int mpsk_splitter_descent_peer(struct mpsk_keyview view, size_t n, uint32_t nw, uint64_t * prefix, const uint64_t * target,
        int nsplit, int level0, int nlevels, uint32_t me, uint32_t p, void * const * boxes, uint32_t seq, uint32_t * err,
        mpsk_stream_t stream)
{
    if (nsplit <= 0 || level0 >= nlevels) return 0;
    if (nw > 16 || nsplit > 63 || p > 64 || me >= p) return (int) cudaErrorInvalidValue;
    auto b = snap((void * const *) boxes, p);
    LAUNCH(stream, K(mpsk_splitter_descent_peer)(view, n, nw, prefix, target, nsplit, level0, nlevels, me, p, b->data(), seq, err, stream));
}

int mpsk_sum_u64(uint64_t * dst, const uint64_t * const * srcs, int nsrc, size_t count, mpsk_stream_t stream)
{
    auto s = snap(srcs, (size_t) (nsrc > 0 ? nsrc : 0));
    LAUNCH(stream, K(mpsk_sum_u64)(dst, s->data(), nsrc, count, stream));
}

int mpsk_merge_samples(const void * recv, size_t elsize, size_t offset, uint32_t width, uint32_t nwords, int is_signed,
        uint32_t p, uint32_t S, uint32_t k, const uint32_t * rdispl, const uint32_t * sstart, uint32_t self_run, const void * self_recv,
        uint64_t * skeys, mpsk_stream_t stream)
{
    auto rd = snap(rdispl, (size_t) p + 1), ss = snap(sstart, (size_t) p + 1);
    LAUNCH(stream, K(mpsk_merge_samples)(recv, elsize, offset, width, nwords, is_signed, p, S, k, rd->data(), ss->data(), self_run, self_recv, skeys, stream));
}

int mpsk_merge_rank_samples(const uint64_t * skeys, uint32_t p, const uint32_t * sstart, uint64_t * sorted_skeys,
        uint32_t * sorted_sid, mpsk_stream_t stream)
{
    if (p > 32) return (int) cudaErrorInvalidValue;
    auto ss = snap(sstart, (size_t) p + 1);
    LAUNCH(stream, K(mpsk_merge_rank_samples)(skeys, p, ss->data(), sorted_skeys, sorted_sid, stream));
}

int mpsk_merge_runs(const void * recv, void * out, size_t elsize, size_t offset, uint32_t width, uint32_t nwords, int is_signed,
        uint32_t p, uint32_t S, uint32_t k, const uint32_t * rdispl, const uint32_t * sstart, uint32_t self_run, const void * self_recv,
        const uint64_t * sorted_skeys, const uint32_t * sorted_sid, uint32_t ntiles, uint32_t * cut, uint32_t * overflow,
        mpsk_stream_t stream)
{
    if (p > 32) return (int) cudaErrorInvalidValue;
    auto rd = snap(rdispl, (size_t) p + 1), ss = snap(sstart, (size_t) p + 1);
    LAUNCH(stream, K(mpsk_merge_runs)(recv, out, elsize, offset, width, nwords, is_signed, p, S, k, rd->data(), ss->data(), self_run, self_recv,
                                      sorted_skeys, sorted_sid, ntiles, cut, overflow, stream));
}

int mpsk_p2p_alltoallv(const void * const * src, void * const * dst, const uint64_t * bytes, const unsigned char * remote,
        int nseg, mpsk_stream_t stream)
{
    const size_t m = (size_t) (nseg > 0 ? nseg : 0);
    auto s = snap(src, m); auto d = snap(dst, m); auto b = snap(bytes, m); auto r = snap(remote, m);
    LAUNCH(stream, K(mpsk_p2p_alltoallv)(s->data(), d->data(), b->data(), r->data(), nseg, stream));
}

int mpsk_p2p_gather_alltoallv(const void * base, const uint32_t * const * idx, void * const * dst, const uint64_t * nrec,
        size_t elsize, int nseg, mpsk_stream_t stream)
{
    const size_t m = (size_t) (nseg > 0 ? nseg : 0);
    auto i = snap(idx, m); auto d = snap(dst, m); auto c = snap(nrec, m);
    LAUNCH(stream, K(mpsk_p2p_gather_alltoallv)(base, i->data(), d->data(), c->data(), elsize, nseg, stream));
}

int mpsk_checksum(const void * base, size_t nbytes, uint64_t * sum, mpsk_stream_t stream)
{ LAUNCH(stream, K(mpsk_checksum)(base, nbytes, sum, stream)); }

int mpsk_multiset_hash(const void * base, size_t n, size_t elsize, uint64_t * out, mpsk_stream_t stream)
{ LAUNCH(stream, K(mpsk_multiset_hash)(base, n, elsize, out, stream)); }

int mpsk_generate(void * dst, size_t n, size_t elsize, int kind, uint64_t seed, uint64_t rank, uint64_t nranks, mpsk_stream_t stream)
{ LAUNCH(stream, K(mpsk_generate)(dst, n, elsize, kind, seed, rank, nranks, stream)); }

int mpsk_check_sorted(const void * base, size_t n, size_t elsize, size_t offset, uint32_t width, uint32_t nwords, int is_signed,
        int check_ties, size_t tie_offset, uint64_t * violations, uint64_t * firstlast, mpsk_stream_t stream)
{ LAUNCH(stream, K(mpsk_check_sorted)(base, n, elsize, offset, width, nwords, is_signed, check_ties, tie_offset, violations, firstlast, stream)); }

}   /* extern "C" */
