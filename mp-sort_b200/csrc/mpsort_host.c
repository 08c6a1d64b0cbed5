/*
 * mpsort_host.c -- C host orchestration of the B200 distributed histogram sort.
 *
 * Mirrors, phase by phase, mpsort_mpi_newarray_impl and mpsort_mpi_histogram_sort
 * of the reference (mpsort-mpi.c:161-331, :333-604); every device step is a
 * hand-written sm_100a kernel from mpsort_kernels.cu, every collective goes
 * through mpsort_comm.c (NCCL over NVLink in production).
 *
 *   reference phase (timer name)      here
 *   -------------------------------   -------------------------------------------
 *   FirstSort   radix_sort :369       local_sort(): extract+hist, onesweep passes
 *   PmaxPmin    :375, :606-661        one host allgather of {n, outn, kmin, kmax}
 *   findP       :383-441              byte-wise descent: count kernel + allreduce
 *                                     + select kernel per level, no host round trip
 *   LayDistr    :450-456              one host allgather of local CLT/CLE rows
 *   LaySolve    :460-464, :663-727    mpsort_solve_layout() on every rank
 *   Exchange    :571-592              payload gather (pack, index mode) + DMA peer copies
 *                                     in two parts (or grouped send/recv)
 *   SecondSort  radix_sort :597       stable p-way merge of the received runs (or re-sort)
 */
#include <stdarg.h>
#include <string.h>

#include "mpsort_internal.h"

void mpsort_cumulative_counts(int p, const int64_t * outnmemb, int64_t * C);
int mpsort_key_range(int p, uint32_t nw, const int64_t * nmemb,
        const uint64_t * kmin, const uint64_t * kmax,
        uint64_t * Pmin, uint64_t * Pmax, uint64_t * prefix);

/* ------------------------------------------------------------------------- */
/* options (reference mpsort-mpi.c:17, :731-767)                              */

static int _mpsort_mpi_options = 0;

static void _mpsort_mpi_parse_env(void)
{
    static int parsed = 0;
    if (parsed) return;
    parsed = 1;
    if (getenv("MPSORT_DISABLE_SPARSE_ALLTOALLV"))
        _mpsort_mpi_options |= MPSORT_DISABLE_SPARSE_ALLTOALLV;
    if (getenv("MPSORT_DISABLE_GATHER_SORT"))
        _mpsort_mpi_options |= MPSORT_DISABLE_GATHER_SORT;
    /* the reference looks up "MPSORT_REQUIRE_GATHER_SORT " with a trailing space
     * (mpsort-mpi.c:741) and so never sees it; we honour the documented name */
    if (getenv("MPSORT_REQUIRE_GATHER_SORT"))
        _mpsort_mpi_options |= MPSORT_REQUIRE_GATHER_SORT;
    if (getenv("MPSORT_REQUIRE_SPARSE_ALLTOALLV"))
        _mpsort_mpi_options |= MPSORT_REQUIRE_SPARSE_ALLTOALLV;
    if (getenv("MPSORT_VERIFY_CHECKSUM"))
        _mpsort_mpi_options |= MPSORT_VERIFY_CHECKSUM;
}

void mpsort_mpi_set_options(int options)
{
    _mpsort_mpi_parse_env();
    _mpsort_mpi_options |= options;
}

int mpsort_mpi_has_options(int options)
{
    _mpsort_mpi_parse_env();
    return _mpsort_mpi_options & options;
}

void mpsort_mpi_unset_options(int options)
{
    _mpsort_mpi_parse_env();
    _mpsort_mpi_options &= ~options;
}

/* ------------------------------------------------------------------------- */
/* timers (reference mpsort-mpi.c:105-127): CUDA events on the comm's stream  */


static struct {
    char name[MPS_MAX_TIMERS][20];
    double seconds[MPS_MAX_TIMERS];
    int n;
} g_last_run;
static pthread_mutex_t g_last_run_lock = PTHREAD_MUTEX_INITIALIZER;

static void timer_reset(struct mpsort_comm * c)
{
    int i;
    struct mps_timers * T = &c->timers;
    if (!T->created) {
        for (i = 0; i < MPS_MAX_TIMERS; i++) CUDA_OK(c, cudaEventCreate(&T->ev[i]));
        T->created = 1;
    }
    T->n = 0;
}

static void timer_mark(struct mpsort_comm * c, const char * name)
{
    struct mps_timers * T = &c->timers;
    if (T->n >= MPS_MAX_TIMERS) return;
    CUDA_OK(c, cudaEventRecord(T->ev[T->n], c->stream));
    snprintf(T->name[T->n], sizeof(T->name[0]), "%s", name);
    T->n++;
}

/* called after the stream was synchronised; rank 0 publishes (the reference
 * broadcasts the leader's timers, mpsort-mpi.c:306-313) */
static void timer_publish(struct mpsort_comm * c)
{
    int i;
    struct mps_timers * T = &c->timers;
    if (c->rank != 0 || T->n < 1) return;
    pthread_mutex_lock(&g_last_run_lock);
    g_last_run.n = 0;
    for (i = 1; i < T->n; i++) {
        float ms = 0;
        if (0 == strcmp(T->name[i], "END")) break;
        if (cudaEventElapsedTime(&ms, T->ev[i - 1], T->ev[i]) != cudaSuccess) ms = 0;
        snprintf(g_last_run.name[g_last_run.n], sizeof(g_last_run.name[0]), "%s", T->name[i]);
        g_last_run.seconds[g_last_run.n] = ms * 1e-3;
        g_last_run.n++;
    }
    pthread_mutex_unlock(&g_last_run_lock);
}

void mpsort_mpi_report_last_run(void)
{
    int i;
    pthread_mutex_lock(&g_last_run_lock);
    for (i = 0; i < g_last_run.n; i++) printf("%s: %g\n", g_last_run.name[i], g_last_run.seconds[i]);
    pthread_mutex_unlock(&g_last_run_lock);
}

int mpsort_mpi_get_last_run(const char ** names, double * seconds, int max)
{
    int i;
    pthread_mutex_lock(&g_last_run_lock);
    for (i = 0; i < g_last_run.n && i < max; i++) {
        if (names) names[i] = g_last_run.name[i];
        if (seconds) seconds[i] = g_last_run.seconds[i];
    }
    i = g_last_run.n;
    pthread_mutex_unlock(&g_last_run_lock);
    return i;
}

void mpsort_comm_last_stats(mpsort_comm_t c, struct mpsort_last_stats * st, int64_t * sendcounts, int max)
{
    int i;
    if (st) *st = c->stats;
    if (sendcounts) for (i = 0; i < c->size && i < max; i++) sendcounts[i] = c->sendcounts[i];
}

/* ------------------------------------------------------------------------- */
/* host buffers in chunks: see struct mpsort_comm                */

/* host buffers of at least MPSORT_CHUNK_MIN_BYTES (64 MiB) move in chunks of MPSORT_CHUNK_BYTES (256 MiB);
 * the tests lower both to run the chunked flow on small arrays */
static size_t io_env_bytes(const char * name, size_t dflt)
{
    const char * e = getenv(name);
    const long long v = e ? atoll(e) : 0;
    return v > 0 ? (size_t) v : dflt;
}
#define MPS_CHUNK_MIN_BYTES io_env_bytes("MPSORT_CHUNK_MIN_BYTES", (size_t) 64 << 20)
#define MPS_CHUNK_BYTES io_env_bytes("MPSORT_CHUNK_BYTES", (size_t) 256 << 20)

static void io_create(struct mpsort_comm * c)
{
    int k;
    if (c->io_created) return;
    CUDA_OK(c, cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
    CUDA_OK(c, cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    for (k = 0; k < MPS_MAX_CHUNKS; k++) CUDA_OK(c, cudaEventCreateWithFlags(&c->in.ev[k], cudaEventDisableTiming));
    for (k = 0; k < 2; k++) CUDA_OK(c, cudaEventCreateWithFlags(&c->io_ev[k], cudaEventDisableTiming));
    c->io_created = 1;
}

static int io_chunking_enabled(void)
{
    static int v = -1;
    if (v < 0) v = getenv("MPSORT_NO_HOST_CHUNKS") ? 0 : 1;
    return v;
}

/* host -> device: one copy on c->stream, or (large inputs) chunks on the copy stream that the readers of
 * `din` wait for range by range (mps_input_wait) */
static void input_stage(struct mpsort_comm * c, void * din, const void * base, size_t nmemb, size_t elsize)
{
    const size_t bytes = nmemb * elsize;
    c->in.dbase = NULL;
    if (bytes < MPS_CHUNK_MIN_BYTES || !io_chunking_enabled()) {
        if (bytes) CUDA_OK(c, cudaMemcpyAsync(din, base, bytes, cudaMemcpyHostToDevice, c->stream));
        return;
    }
    io_create(c);
    size_t nch = (bytes + MPS_CHUNK_BYTES - 1) / MPS_CHUNK_BYTES;
    if (nch > MPS_MAX_CHUNKS) nch = MPS_MAX_CHUNKS;
    /* whole fix-up / merge tiles per chunk is not needed on the way in: any record boundary will do,
     * kept at a multiple of 4096 records so that vector loads of every kernel stay aligned */
    size_t per = ((nmemb + nch - 1) / nch + 4095) & ~(size_t) 4095;
    nch = (nmemb + per - 1) / per;
    c->in.dbase = din; c->in.n = nmemb; c->in.elsize = elsize; c->in.chunk = per; c->in.nchunks = (int) nch; c->in.waited = 0;
    /* the copy stream starts after whatever c->stream has queued on `din` (an earlier call's readers) */
    CUDA_OK(c, cudaEventRecord(c->io_ev[0], c->stream));
    CUDA_OK(c, cudaStreamWaitEvent(c->h2d_stream, c->io_ev[0], 0));
    size_t k;
    for (k = 0; k < nch; k++) {
        const size_t a = k * per, b = (a + per < nmemb) ? a + per : nmemb;
        CUDA_OK(c, cudaMemcpyAsync((char *) din + a * elsize, (const char *) base + a * elsize, (b - a) * elsize,
                                   cudaMemcpyHostToDevice, c->h2d_stream));
        CUDA_OK(c, cudaEventRecord(c->in.ev[k], c->h2d_stream));
    }
}

void mps_input_wait(struct mpsort_comm * c, const void * dbase, size_t first, size_t count)
{
    if (!c->in.dbase || dbase != c->in.dbase || count == 0) return;
    size_t last = (first + count - 1) / c->in.chunk;
    if (last >= (size_t) c->in.nchunks) last = (size_t) c->in.nchunks - 1;
    if ((int) last < c->in.waited) return;
    /* the copy stream runs in order: waiting for chunk `last` is waiting for all before it */
    CUDA_OK(c, cudaStreamWaitEvent(c->stream, c->in.ev[last], 0));
    c->in.waited = (int) last + 1;
}

int mps_input_ranges(struct mpsort_comm * c, const void * dbase, size_t n, size_t * a)
{
    int k;
    a[0] = 0;
    if (!c->in.dbase || dbase != c->in.dbase || n != c->in.n) { a[1] = n; return 1; }
    for (k = 0; k < c->in.nchunks; k++) a[k + 1] = ((size_t) (k + 1) * c->in.chunk < n) ? (size_t) (k + 1) * c->in.chunk : n;
    return c->in.nchunks;
}

static void output_begin(struct mpsort_comm * c, void * host, const void * dout, size_t outnmemb, size_t elsize)
{
    c->outp.host = NULL;
    if (!host || outnmemb * elsize < MPS_CHUNK_MIN_BYTES || !io_chunking_enabled()) return;
    io_create(c);
    c->outp.host = host; c->outp.dout = dout; c->outp.elsize = elsize; c->outp.total = outnmemb; c->outp.done = 0;
}

void mps_output_ready(struct mpsort_comm * c, size_t upto)
{
    if (!c->outp.host) return;
    if (upto > c->outp.total) upto = c->outp.total;
    if (upto <= c->outp.done) return;
    /* the copy stream waits for the producer (c->stream may be the merge stream right now) */
    CUDA_OK(c, cudaEventRecord(c->io_ev[1], c->stream));
    CUDA_OK(c, cudaStreamWaitEvent(c->d2h_stream, c->io_ev[1], 0));
    CUDA_OK(c, cudaMemcpyAsync((char *) c->outp.host + c->outp.done * c->outp.elsize,
                               (const char *) c->outp.dout + c->outp.done * c->outp.elsize,
                               (upto - c->outp.done) * c->outp.elsize, cudaMemcpyDeviceToHost, c->d2h_stream));
    c->outp.done = upto;
}

void mps_output_reset(struct mpsort_comm * c)
{
    if (c->outp.host) c->outp.done = 0;
}

/* everything not yet on its way leaves now; c->stream then waits for the copy stream */
static void output_finish(struct mpsort_comm * c)
{
    if (!c->outp.host) return;
    mps_output_ready(c, c->outp.total);
    CUDA_OK(c, cudaEventRecord(c->io_ev[1], c->d2h_stream));
    CUDA_OK(c, cudaStreamWaitEvent(c->stream, c->io_ev[1], 0));
    c->outp.host = NULL;
}

/* ------------------------------------------------------------------------- */
/* local sort: replaces radix_sort (radixsort.c:35-44)                        */

struct sorted_view {
    const uint64_t * skeys;   /* sorted packed key words, word w at skeys + w*stride */
    size_t stride;
    const uint32_t * idx;     /* idx[i] = original position of the i-th smallest record */
    uint32_t nw;
    uint32_t npasses;
    struct mpsk_keyview kv;   /* how the splitter kernels read the sorted keys */
    const void * sorted_recs; /* record mode: the records themselves, already in sorted order */
};

static uint32_t key_words(const struct mpsort_radix_desc * d)
{
    return (uint32_t) (((size_t) d->width * d->nwords + 7) / 8);
}

/*
 * Stable LSD radix sort of the n records at device pointer `dbase` by their
 * descriptor key. Produces the permutation and (when want_keys) the sorted packed
 * keys. Multi-word keys are sorted one 64-bit word at a time from the least
 * significant word: each word costs 8 (key,idx) passes, and the next word is
 * brought into the current order with one 8-byte gather.
 * Digits that are constant over the whole array are skipped (their pass would be
 * the identity permutation): small ids, 4-byte keys and post-exchange key ranges
 * all profit.
 */
#define MPS_REBASE_MIN_ITEMS ((size_t) 1 << 20)

static void local_sort(struct mpsort_comm * c, const void * dbase, size_t n, size_t elsize,
        const struct mpsort_radix_desc * desc, int want_keys, int allow_rebase, struct sorted_view * out)
{
    const uint32_t nw = key_words(desc);
    uint32_t g, d;
    memset(out, 0, sizeof(*out));
    out->nw = nw;
    out->stride = n;
    if (n == 0) return;
    if (n > MPSK_MAX_ITEMS)
        mps_fatal(c, __FILE__, __LINE__, "%zu local items exceed the supported maximum %zu per rank", n, (size_t) MPSK_MAX_ITEMS);

    uint64_t * kw = (uint64_t *) mps_arena_get(c, MPS_S_KW, (size_t) nw * n * sizeof(uint64_t));
    uint64_t * kb = (uint64_t *) mps_arena_get(c, MPS_S_KB, n * sizeof(uint64_t));
    uint64_t * ka = nw > 1 ? (uint64_t *) mps_arena_get(c, MPS_S_KA, n * sizeof(uint64_t)) : kw;
    uint32_t * ia = (uint32_t *) mps_arena_get(c, MPS_S_IA, n * sizeof(uint32_t));
    uint32_t * ib = (uint32_t *) mps_arena_get(c, MPS_S_IB, n * sizeof(uint32_t));
    const size_t nhist = (size_t) nw * 8;
    uint32_t * hist = (uint32_t *) mps_arena_get(c, MPS_S_HIST, nhist * 256 * sizeof(uint32_t) * 2 + 64);
    uint32_t * bins = hist + nhist * 256;
    uint64_t * minmax = (uint64_t *) (bins + nhist * 256);      /* {min, max} of single-word keys */
    void * scratch = mps_arena_get(c, MPS_S_SCRATCH, mpsk_onesweep_scratch_bytes(n));
    uint64_t add = 0;

    CUDA_OK(c, cudaMemsetAsync(hist, 0, nhist * 256 * sizeof(uint32_t), c->stream));
    {
        uint64_t * hmm = (uint64_t *) mps_host_stage(c, nhist * 256 * sizeof(uint32_t) + 64) + (nhist * 256 * sizeof(uint32_t)) / 8;
        hmm[0] = ~0ULL; hmm[1] = 0;
        CUDA_OK(c, cudaMemcpyAsync(minmax, hmm, 2 * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    }
    {
        /* a host input that is still arriving is read chunk by chunk (histograms and min/max accumulate) */
        size_t ra[MPS_MAX_CHUNKS + 1];
        const int nr = mps_input_ranges(c, dbase, n, ra);
        int i;
        for (i = 0; i < nr; i++) {
            mps_input_wait(c, dbase, ra[i], ra[i + 1] - ra[i]);
            for (g = 0; g < nw; g++)
                KERN_T(c, MPS_K_EXTRACT, mpsk_extract_keys((const char *) dbase + ra[i] * elsize, ra[i + 1] - ra[i], elsize, desc->offset,
                                             desc->width, desc->nwords, desc->is_signed, g, 0, kw + (size_t) g * n + ra[i],
                                             hist + (size_t) g * 8 * 256, (nw == 1 ? minmax : NULL), c->stream));
        }
    }
    KERN_T(c, MPS_K_EXTRACT, mpsk_scan_histograms(hist, bins, (int) nhist, c->stream));

    /* which digits are constant? (one small D2H per sort; bins and min/max ride along) */
    uint32_t * hhist = (uint32_t *) mps_host_stage(c, nhist * 256 * sizeof(uint32_t) + 64);
    unsigned char skip[MPS_MAX_KEY_WORDS * 8];
    uint32_t todo = 0;
    int round;
    for (round = 0; round < 2; round++) {
        CUDA_OK(c, cudaMemcpyAsync(hhist, hist, nhist * 256 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(c, cudaMemcpyAsync((char *) hhist + nhist * 256 * sizeof(uint32_t), minmax, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(c, cudaStreamSynchronize(c->stream));
        todo = 0;
        for (g = 0; g < nhist; g++) {
            uint32_t b;
            skip[g] = 0;
            for (b = 0; b < 256; b++) {
                if (hhist[(size_t) g * 256 + b] == (uint32_t) n) { skip[g] = 1; break; }
            }
            if (!skip[g]) todo++;
        }
        if (round == 1 || !allow_rebase || nw != 1 || n < MPS_REBASE_MIN_ITEMS || getenv("MPSORT_NO_REBASE")) break;
        /* Range compression: keys that sit in a narrow range far from zero (small signed
         * ids are 0x7fff.. / 0x8000.. after the sign flip) vary in every byte, yet
         * key - min needs only ceil(log256(max - min)) passes. One extra 16 B/key pass
         * rewrites the keys relative to the minimum and recounts the digits. */
        const uint64_t * mm = (const uint64_t *) ((char *) hhist + nhist * 256 * sizeof(uint32_t));
        const uint64_t kmin = mm[0], range = mm[1] - mm[0];
        uint32_t need = 0;
        while (need < 8 && (range >> (8 * need)) != 0) need++;
        if (need + 1 > todo) break;                 /* saves less than one pass */
        add = kmin;
        CUDA_OK(c, cudaMemsetAsync(hist, 0, nhist * 256 * sizeof(uint32_t), c->stream));
        KERN_T(c, MPS_K_EXTRACT, mpsk_extract_keys(kw, n, 8, 0, 8, 1, 0, 0, kmin, kw, hist, NULL, c->stream));
        KERN_T(c, MPS_K_EXTRACT, mpsk_scan_histograms(hist, bins, (int) nhist, c->stream));
        c->stats.rebased = 1;
    }
    if (todo == 0) { skip[0] = 0; todo = 1; }  /* all keys equal: one pass yields the identity idx */

    const uint32_t * cur_idx = NULL;     /* NULL = identity */
    const uint64_t * cur_keys = NULL;    /* sorted keys of the word in progress */
    uint32_t npasses = 0;
    for (g = 0; g < nw; g++) {
        int any = 0;
        for (d = 0; d < 8; d++) if (!skip[g * 8 + d]) any = 1;
        if (!any) continue;
        /* keys of word g in the current order */
        const uint64_t * kin;
        uint64_t * kout;
        if (cur_idx == NULL) {
            kin = kw + (size_t) g * n;
        } else {
            KERN_T(c, MPS_K_GATHER_KEYS, mpsk_gather_u64(kw + (size_t) g * n, cur_idx, ka == kw ? kb : ka, n, c->stream));
            kin = (ka == kw ? kb : ka);
        }
        for (d = 0; d < 8; d++) {
            if (skip[g * 8 + d]) continue;
            /* ping-pong: never write into the original-order words of a multi-word key */
            if (nw == 1) kout = (kin == kw) ? kb : kw;
            else kout = (kin == ka) ? kb : ka;
            uint32_t * vout = (cur_idx == ia) ? ib : ia;
            KERN_T(c, MPS_K_ONESWEEP, mpsk_onesweep_pass(kin, cur_idx, kout, vout, n, (int) (8 * d),
                                          bins + (size_t) (g * 8 + d) * 256, scratch, c->stream));
            kin = kout;
            cur_idx = vout;
            npasses++;
        }
        cur_keys = kin;
    }
    out->idx = cur_idx;
    out->npasses = npasses;
    if (want_keys) {
        if (nw == 1) {
            out->skeys = cur_keys;
            out->kv.base = cur_keys; out->kv.item_stride = 8; out->kv.word_stride = 0; out->kv.flip = 0; out->kv.add = add;
        } else {
            uint64_t * sk = (uint64_t *) mps_arena_get(c, MPS_S_SK, (size_t) nw * n * sizeof(uint64_t));
            for (g = 0; g < nw; g++)
                KERN_T(c, MPS_K_GATHER_KEYS, mpsk_gather_u64(kw + (size_t) g * n, cur_idx, sk + (size_t) g * n, n, c->stream));
            out->skeys = sk;
            out->kv.base = sk; out->kv.item_stride = 8; out->kv.word_stride = n * sizeof(uint64_t); out->kv.flip = 0; out->kv.add = 0;
        }
    }
}

/*
 * Record mode: a 16-byte record that is nothing but an aligned 8-byte key and 8 more
 * bytes is carried through the passes itself (mpsk_onesweep_pass_rec16): no key
 * extraction, no index array, no payload gather. `dest` receives the sorted records;
 * it may equal dbase (in place). dbase is only read when dest != dbase.
 */
/* (the shape alone: aligned buffers decide the rest) */
static int rec16_shape(size_t elsize, const struct mpsort_radix_desc * d)
{
    if (getenv("MPSORT_NO_REC16")) return 0;
    if (d->nwords != 1 || d->width != 8) return 0;
    return (elsize == 16 && (d->offset == 0 || d->offset == 8)) || (elsize == 8 && d->offset == 0);
}

static int rec16_applicable(size_t elsize, const struct mpsort_radix_desc * d, const void * dbase, const void * dest)
{
    if (getenv("MPSORT_NO_REC16")) return 0;
    if (d->nwords != 1 || d->width != 8) return 0;
    if (elsize == 16) return (d->offset == 0 || d->offset == 8) && ((((uintptr_t) dbase) | ((uintptr_t) dest)) & 15) == 0;
    if (elsize == 8) return d->offset == 0 && ((((uintptr_t) dbase) | ((uintptr_t) dest)) & 7) == 0;   /* bare keys */
    return 0;
}

/* P stable passes over the given digits, source `src` (read only unless it is also
 * the destination), result in `dest`; `Y` and the KB slot are the ping-pong temps */
static void rec16_passes(struct mpsort_comm * c, const void * src, size_t n, size_t E, const struct mpsort_radix_desc * desc,
        void * dest, const int * digits, int P, const uint32_t * bins, void * scratch)
{
    const uint64_t flip = desc->is_signed ? (1ULL << 63) : 0ULL;
    void * Y = mps_arena_get(c, MPS_S_KW, n * E);
    int i;
    for (i = 0; i < P; i++) {
        const int remaining_after = P - 1 - i;
        void * tgt = (remaining_after % 2 == 0) ? dest : Y;
        if (tgt == src) tgt = mps_arena_get(c, MPS_S_KB, n * E);   /* first pass of an odd in-place chain */
        KERN_T(c, MPS_K_ONESWEEP_REC, mpsk_onesweep_pass_rec(src, tgt, n, E, 8 * digits[i], desc->offset == 8, flip,
                                                              bins + (size_t) digits[i] * 256, scratch, c->stream));
        src = tgt;
    }
}

#define MPS_HYBRID_MIN_ITEMS ((size_t) 1 << 22)
#define MPS_HYBRID_SAMPLES 65536u
#define MPS_HYBRID_MAX_LONG_RUNS 64u

static void local_sort(struct mpsort_comm * c, const void * dbase, size_t n, size_t elsize,
        const struct mpsort_radix_desc * desc, int want_keys, int allow_rebase, struct sorted_view * out);

/* The predictor of the hybrid sort. Does the high part (key >> lobits) look nearly distinct?
 * mpsk_prefix_pairs counts the equal PAIRS c among the high parts of 65536 evenly spaced
 * records (a hash table in L2, one launch, for up to two values of lobits at once):
 * E[c] = s^2/(2n) * (mean length of the run a random record sits in, minus one).
 * The run fix-up costs about that many comparisons per record. */
#define MPS_PRED_LOG2_TABLE 18u
struct predictor { uint32_t lobits[2]; int nl; uint64_t * d_pairs; };

static void predictor_launch(struct mpsort_comm * c, const void * dbase, size_t n, size_t E,
        const struct mpsort_radix_desc * desc, struct predictor * pr)
{
    const uint64_t flip = desc->is_signed ? (1ULL << 63) : 0ULL;
    const size_t tbytes = (size_t) (pr->nl + 1) * 2 * ((size_t) 1 << MPS_PRED_LOG2_TABLE) * sizeof(uint64_t);
    uint64_t * tab = (uint64_t *) mps_arena_get(c, MPS_S_PRED, tbytes + 64);
    pr->d_pairs = tab + tbytes / sizeof(uint64_t);
    CUDA_OK(c, cudaMemsetAsync(tab, 0, tbytes + 64, c->stream));
    KERN_T(c, MPS_K_HYBRID, mpsk_prefix_pairs(dbase, n, E, MPS_HYBRID_SAMPLES, desc->offset == 8, flip, pr->lobits, (uint32_t) pr->nl,
                                              tab, MPS_PRED_LOG2_TABLE, pr->d_pairs, c->stream));
}

/* pairs -> estimated mean run length */
static double predictor_mean_run(uint64_t pairs, size_t n)
{
    const double s = (double) MPS_HYBRID_SAMPLES;
    return 1.0 + (double) pairs * 2.0 * (double) n / (s * s);
}

static void local_sort_rec16(struct mpsort_comm * c, const void * dbase, size_t n, size_t E,
        const struct mpsort_radix_desc * desc, void * dest, struct sorted_view * out)
{
    const uint64_t flip = desc->is_signed ? (1ULL << 63) : 0ULL;
    uint32_t d;
    memset(out, 0, sizeof(*out));
    out->nw = 1;
    out->stride = n;
    out->sorted_recs = dest;
    out->kv.base = (const char *) dest + desc->offset;
    out->kv.item_stride = E; out->kv.word_stride = 0; out->kv.flip = flip; out->kv.add = 0;
    if (n == 0) return;
    if (n > MPSK_MAX_ITEMS)
        mps_fatal(c, __FILE__, __LINE__, "%zu local items exceed the supported maximum %zu per rank", n, (size_t) MPSK_MAX_ITEMS);

    uint32_t * hist = (uint32_t *) mps_arena_get(c, MPS_S_HIST, 8 * 256 * sizeof(uint32_t) * 2 + 64);
    uint32_t * bins = hist + 8 * 256;
    uint64_t * ddiff = (uint64_t *) (bins + 8 * 256);        /* OR of key ^ key[0]: which bytes vary */
    void * scratch = mps_arena_get(c, MPS_S_SCRATCH, mpsk_onesweep_scratch_bytes(n));
    const int khi = desc->offset == 8;
    const int hybrid_ok = n >= MPS_HYBRID_MIN_ITEMS && !getenv("MPSORT_NO_HYBRID");
    CUDA_OK(c, cudaMemsetAsync(hist, 0, 8 * 256 * sizeof(uint32_t), c->stream));
    CUDA_OK(c, cudaMemsetAsync(ddiff, 0, sizeof(uint64_t), c->stream));
    /* The hybrid sort only needs the counts of the four (or five) most significant digits, and
     * counting four digits instead of eight makes the histogram pass HBM-bound. A preview over
     * 4096 evenly spaced keys says whether that is where this input is heading (bytes that vary
     * in the sample vary in the whole array) and which high parts the predictor should look at:
     * the predictor then runs beside the histogram pass and both are read back together --
     * two host round trips before the passes instead of five. */
    int have_low = 1;
    struct predictor pr;
    memset(&pr, 0, sizeof(pr));
    size_t ra[MPS_MAX_CHUNKS + 1];
    const int nr = mps_input_ranges(c, dbase, n, ra);
    int ri;
    if (hybrid_ok && !getenv("MPSORT_NO_HIST4")) {
        uint64_t * hd = (uint64_t *) mps_host_stage(c, sizeof(uint64_t));
        /* (of a host input that is still arriving, the preview looks at the first chunk) */
        mps_input_wait(c, dbase, 0, ra[1]);
        KERN_T(c, MPS_K_EXTRACT, mpsk_rec_sample_diff(dbase, ra[1], E, khi, 4096, ddiff, c->stream));
        CUDA_OK(c, cudaMemcpyAsync(hd, ddiff, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(c, cudaMemsetAsync(ddiff, 0, sizeof(uint64_t), c->stream));
        CUDA_OK(c, cudaStreamSynchronize(c->stream));
        int vary = 0, top4 = 1, pd[8];
        for (d = 0; d < 8; d++) {
            const int v = ((*hd >> (8 * d)) & 255u) != 0;
            if (v) pd[vary++] = (int) d;
            if (d >= 4 && !v) top4 = 0;
        }
        if (top4 && vary >= 6) have_low = 0;
        if (vary >= 6) {
            /* the high parts the exact digit list will most likely ask about */
            pr.lobits[pr.nl++] = 8u * (uint32_t) pd[vary - 4];
            if (vary >= 7) pr.lobits[pr.nl++] = 8u * (uint32_t) pd[vary - 5];
        }
    }
    for (ri = 0; ri < nr; ri++) {
        /* chunk by chunk behind the copy of a host input; one launch otherwise */
        const char * part = (const char *) dbase + ra[ri] * E;
        mps_input_wait(c, dbase, ra[ri], ra[ri + 1] - ra[ri]);
        if (have_low) KERN_T(c, MPS_K_EXTRACT, mpsk_rec_histograms(part, ra[ri + 1] - ra[ri], E, khi, flip, 0, 8, hist, ddiff, dbase, c->stream));
        else KERN_T(c, MPS_K_EXTRACT, mpsk_rec_histograms(part, ra[ri + 1] - ra[ri], E, khi, flip, 4, 4, hist, ddiff, dbase, c->stream));
    }
    KERN_T(c, MPS_K_EXTRACT, mpsk_scan_histograms(hist, bins, 8, c->stream));
    if (pr.nl) predictor_launch(c, dbase, n, E, desc, &pr);
    const size_t hbytes = 8 * 256 * sizeof(uint32_t);
    uint32_t * hhist = (uint32_t *) mps_host_stage(c, hbytes + 8 * sizeof(uint64_t));
    uint64_t * htail = (uint64_t *) ((char *) hhist + hbytes);      /* [0] diff, [1..3] predictor pairs */
    CUDA_OK(c, cudaMemcpyAsync(hhist, hist, hbytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(htail, ddiff, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    if (pr.nl) CUDA_OK(c, cudaMemcpyAsync(htail + 1, pr.d_pairs, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    int digits[8], P = 0;
    /* (pairs of samples that drew the same record are in every count: off they go) */
    uint64_t pairs_of[2] = { 0, 0 };
    if (pr.nl) {
        const uint64_t dup = htail[1 + pr.nl];
        pairs_of[0] = htail[1] - dup;
        if (pr.nl > 1) pairs_of[1] = htail[2] - dup;
    }
    {
        const uint64_t diff = htail[0];
        for (d = 0; d < 8; d++) if ((diff >> (8 * d)) & 255u) digits[P++] = (int) d;
    }
    out->npasses = (uint32_t) P;
    if (P == 0) {
        if (dest != dbase) CUDA_OK(c, cudaMemcpyAsync(dest, dbase, n * E, cudaMemcpyDeviceToDevice, c->stream));
        return;
    }
/* counts of the four low digits, when the first histogram pass left them out */
#define ENSURE_LOW_HISTOGRAMS() do { if (!have_low) { \
        KERN_T(c, MPS_K_EXTRACT, mpsk_rec_histograms(dbase, n, E, khi, flip, 0, 4, hist, NULL, NULL, c->stream)); \
        KERN_T(c, MPS_K_EXTRACT, mpsk_scan_histograms(hist, bins, 8, c->stream)); \
        have_low = 1; } } while (0)

    /* ---- hybrid: passes over the four (or five) most significant digits + run fix-up */
    if (P >= 6 && hybrid_ok) {
        /* mean run length left by the top four / five digits: from the speculated launch when it asked
         * the right question, else now (one more round trip; only when the preview missed a varying byte) */
        uint32_t want[2] = { 8u * (uint32_t) digits[P - 4], P >= 7 ? 8u * (uint32_t) digits[P - 5] : 0u };
        const int nwant = P >= 7 ? 2 : 1;
        if (pr.nl < nwant || pr.lobits[0] != want[0] || (nwant == 2 && pr.lobits[1] != want[1])) {
            pr.nl = nwant; pr.lobits[0] = want[0]; pr.lobits[1] = want[1];
            predictor_launch(c, dbase, n, E, desc, &pr);
            uint64_t * h = (uint64_t *) mps_host_stage(c, hbytes + 8 * sizeof(uint64_t)) + hbytes / sizeof(uint64_t);
            CUDA_OK(c, cudaMemcpyAsync(h + 1, pr.d_pairs, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
            CUDA_OK(c, cudaStreamSynchronize(c->stream));
            pairs_of[0] = h[1] - h[1 + pr.nl];
            pairs_of[1] = pr.nl > 1 ? h[2] - h[1 + pr.nl] : 0;
        }
        const double run4 = predictor_mean_run(pairs_of[0], n);
        const double run5 = nwant == 2 ? predictor_mean_run(pairs_of[1], n) : 1e30;
        /* A pass costs about what the fix-up of runs nine records longer costs (1.8 ms per pass; fix-up
         * 1.0 ms for singleton runs, 4.0 ms for runs of 16: profiles/r02_call2_static_prefetch_fixup_merge.log):
         * a fifth pass pays when it turns long runs into (nearly) singletons -- mostly sorted ids. */
        int H = 4;
        if (nwant == 2 && run4 > 10.0 && run5 <= 2.0 && !getenv("MPSORT_NO_HYBRID5")) H = 5;
        const double run = H == 4 ? run4 : run5;
        const int yes = run <= 17.0;
        const uint32_t lobits = 8u * (uint32_t) digits[P - H];
        if (yes) {
            if (digits[P - H] < 4) ENSURE_LOW_HISTOGRAMS();
            rec16_passes(c, dbase, n, E, desc, dest, digits + (P - H), H, bins, scratch);
            uint32_t * wl = (uint32_t *) mps_arena_get(c, MPS_S_MERGE_CUT, (2 * MPS_HYBRID_MAX_LONG_RUNS + 64) * sizeof(uint32_t));
            uint32_t * nwork = wl + 2 * MPS_HYBRID_MAX_LONG_RUNS;
            CUDA_OK(c, cudaMemsetAsync(nwork, 0, sizeof(uint32_t), c->stream));
            if (c->outp.host && dest == c->outp.dout && n == c->outp.total) {
                /* the sorted records are a host output: fix up a range of tiles, send it on, next range */
                const size_t tile = mpsk_fixup_tile_items(), ntile = (n + tile - 1) / tile;
                const size_t per = (MPS_CHUNK_BYTES / E + tile - 1) / tile;
                size_t t0;
                for (t0 = 0; t0 < ntile; t0 += per) {
                    KERN_T(c, MPS_K_HYBRID, mpsk_fixup_rec(dest, n, E, desc->offset == 8, flip, lobits, wl, nwork, MPS_HYBRID_MAX_LONG_RUNS,
                                                           t0, per, c->stream));
                    if (t0 + per < ntile) mps_output_ready(c, (t0 + per) * tile);
                }
            } else {
                KERN_T(c, MPS_K_HYBRID, mpsk_fixup_rec(dest, n, E, desc->offset == 8, flip, lobits, wl, nwork, MPS_HYBRID_MAX_LONG_RUNS,
                                                       0, 0, c->stream));
            }
            uint32_t * h = (uint32_t *) mps_host_stage(c, (2 * MPS_HYBRID_MAX_LONG_RUNS + 1) * sizeof(uint32_t));
            CUDA_OK(c, cudaMemcpyAsync(h, nwork, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
            CUDA_OK(c, cudaStreamSynchronize(c->stream));
            const uint32_t nlong = h[0];
            out->npasses = (uint32_t) H;
            c->stats.hybrid = 1;
            c->stats.hybrid_long_runs = nlong;
            if (nlong > 0) mps_output_reset(c);          /* long runs are sorted below: ranges already sent go again */
            if (nlong > MPS_HYBRID_MAX_LONG_RUNS) {
                /* the predictor was wrong: finish with a full stable LSD of what we have
                 * (a permutation of the input in which equal keys kept their order) */
                if (!have_low) {
                    /* (histograms are permutation invariant: counting dest is counting dbase,
                     * which an in-place sort has already overwritten) */
                    KERN_T(c, MPS_K_EXTRACT, mpsk_rec_histograms(dest, n, E, khi, flip, 0, 4, hist, NULL, NULL, c->stream));
                    KERN_T(c, MPS_K_EXTRACT, mpsk_scan_histograms(hist, bins, 8, c->stream));
                    have_low = 1;
                }
                rec16_passes(c, dest, n, E, desc, dest, digits, P, bins, scratch);
                out->npasses = (uint32_t) H + (uint32_t) P;
            } else if (nlong > 0) {
                uint32_t starts[MPS_HYBRID_MAX_LONG_RUNS], lens[MPS_HYBRID_MAX_LONG_RUNS], e;
                KERN_T(c, MPS_K_HYBRID, mpsk_fixup_extents(dest, n, E, desc->offset == 8, flip, lobits, wl, nlong, wl + MPS_HYBRID_MAX_LONG_RUNS, c->stream));
                CUDA_OK(c, cudaMemcpyAsync(h, wl, 2 * MPS_HYBRID_MAX_LONG_RUNS * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
                CUDA_OK(c, cudaStreamSynchronize(c->stream));
                for (e = 0; e < nlong; e++) { starts[e] = h[e]; lens[e] = h[MPS_HYBRID_MAX_LONG_RUNS + e]; }
                for (e = 0; e < nlong; e++) {
                    /* a long run is a contiguous array whose high digits are constant: an
                     * ordinary in-place record sort of it runs the low passes only */
                    struct sorted_view sub;
                    char * p0 = (char *) dest + (size_t) starts[e] * E;
                    local_sort_rec16(c, p0, lens[e], E, desc, p0, &sub);
                }
            }
            return;
        }
    }
    ENSURE_LOW_HISTOGRAMS();
    rec16_passes(c, dbase, n, E, desc, dest, digits, P, bins, scratch);
#undef ENSURE_LOW_HISTOGRAMS
}

/* ------------------------------------------------------------------------- */
/* SecondSort as a stable p-way merge of the received runs                    */

/* The receive buffer is p sorted runs in source-rank order (reference RecvDispl
 * layout, mpsort-mpi.c:490-503); the reference re-sorts it with the stable merge
 * sort (:597), which leaves equal keys in (source rank, source index) order. A
 * stable p-way merge that prefers the lower run on ties gives the same bytes with
 * E read + E write per record instead of a full radix sort.
 * Returns 0 if the merge ran, 1 if the caller must use the radix path. */
/* will these runs be merged (1) or re-sorted by the radix path (0)? */
static int merge_applicable(int p, size_t outn, const struct mpsort_radix_desc * desc)
{
    if (key_words(desc) != 1 || p > 32 || p < 2 || outn < 4 * mpsk_merge_tile_items() || outn > 0xfffffff0u) return 0;
    return getenv("MPSORT_NO_MERGE") == NULL;
}

/* self_run >= 0: run self_run is not in recvbuf -- its first record is at self_ptr (the rank's own slice,
 * still in the send buffer where the local sort left it; no transport copied it, mps_comm_exchange with
 * p2p.skip_self). The kernels index that run from the base that puts record rdispl[self_run] at self_ptr. */
static int merge_received_runs(struct mpsort_comm * c, int p, const void * recvbuf, const int64_t * rdispl,
        void * dout, size_t outn, size_t elsize, const struct mpsort_radix_desc * desc, int self_run, const void * self_ptr)
{
    const void * self_recv = NULL;
    if (self_run >= 0 && self_ptr)
        self_recv = (const void *) ((uintptr_t) self_ptr - (uintptr_t) ((uint64_t) rdispl[self_run] * elsize));
    const size_t T = mpsk_merge_tile_items_for((const void *) ((uintptr_t) recvbuf | (uintptr_t) self_recv), dout, elsize,
                                               desc->offset, desc->width, desc->nwords, (uint32_t) p);
    int r;
    if (!merge_applicable(p, outn, desc)) return 1;
    /* (k + p) * S <= T with k = 3p: S = T / (4p) rounded down to a power of two */
    uint32_t S = 1;
    while ((size_t) S * 2 * 4 * (size_t) p <= T) S *= 2;
    const uint32_t k = (uint32_t) (T / S) - (uint32_t) p;
    uint32_t rd[33], ss[33];
    ss[0] = 0;
    for (r = 0; r < p; r++) {
        rd[r] = (uint32_t) rdispl[r];
        ss[r + 1] = ss[r] + (uint32_t) ((rdispl[r + 1] - rdispl[r]) / S);
    }
    rd[p] = (uint32_t) rdispl[p];
    const uint32_t ns = ss[p];
    const uint32_t ntiles = ns == 0 ? 1 : (ns + k - 1) / k;

    uint64_t * skeys = (uint64_t *) mps_arena_get(c, MPS_S_MERGE_SAMP, (size_t) (ns ? ns : 1) * sizeof(uint64_t));
    uint64_t * sorted_skeys = (uint64_t *) mps_arena_get(c, MPS_S_MERGE_SORTED, (size_t) (ns ? ns : 1) * sizeof(uint64_t));
    uint32_t * sorted_sid = (uint32_t *) mps_arena_get(c, MPS_S_MERGE_SID, (size_t) (ns ? ns : 1) * sizeof(uint32_t));
    uint32_t * cut = (uint32_t *) mps_arena_get(c, MPS_S_MERGE_CUT, ((size_t) ntiles + 1) * p * sizeof(uint32_t) + 256);
    /* three launches and no host round trip: samples -> their merged order (sample keys are packed,
     * i.e. sign-flipped, already; every run's samples are sorted, so each one ranks itself by binary
     * searches in the other runs' lists) -> tile bounds + tiles. Tiles that break their bound (never
     * happens) are counted in c->d_merge_ovf and looked at once, after the sort's final synchronisation. */
    KERN_T(c, MPS_K_MERGE, mpsk_merge_samples(recvbuf, elsize, desc->offset, desc->width, desc->nwords, desc->is_signed,
                                              (uint32_t) p, S, k, rd, ss, (uint32_t) self_run, self_recv, skeys, c->stream));
    KERN_T(c, MPS_K_MERGE, mpsk_merge_rank_samples(skeys, (uint32_t) p, ss, sorted_skeys, sorted_sid, c->stream));
    KERN_T(c, MPS_K_MERGE, mpsk_merge_runs(recvbuf, dout, elsize, desc->offset, desc->width, desc->nwords, desc->is_signed,
                                           (uint32_t) p, S, k, rd, ss, (uint32_t) self_run, self_recv, sorted_skeys, sorted_sid, ntiles, cut,
                                           c->d_merge_ovf, c->stream));
    c->stats.second_sort_merge_tiles += ntiles;
    return 0;
}

/* bench / test support (include/mpsort_util.h): the SecondSort merge alone, on p sorted runs
 * that already sit in device memory -- one GPU can then time and profile the merge of the
 * 8-GPU configuration at full size. Blocking. Returns 0 if the merge ran, 1 if these runs
 * would take the radix fallback. */
int mpsort_util_merge_runs(mpsort_comm_t c, int p, const void * runs, const int64_t * rdispl,
        void * out, size_t elsize, const struct mpsort_radix_desc * desc)
{
    CUDA_OK(c, cudaSetDevice(c->device));
    mps_merge_ovf_begin(c);
    const int rc = merge_received_runs(c, p, runs, rdispl, out, (size_t) (rdispl[p] - rdispl[0]), elsize, desc, -1, NULL);
    mps_merge_ovf_fetch(c);
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    mps_merge_ovf_check(c);
    mps_kt_collect(c);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* helpers                                                                    */

static int is_device_pointer(struct mpsort_comm * c, const void * p)
{
    struct cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    (void) c;
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static uint64_t device_checksum(struct mpsort_comm * c, const void * d, size_t nbytes)
{
    uint64_t * dsum = (uint64_t *) mps_arena_get(c, MPS_S_MISC, 256);
    uint64_t h = 0, all[MPS_MAX_RANKS];
    int j;
    CUDA_OK(c, cudaMemsetAsync(dsum, 0, sizeof(uint64_t), c->stream));
    KERN_T(c, MPS_K_CHECKSUM, mpsk_checksum(d, nbytes, dsum, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(&h, dsum, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    mpsort_comm_allgather_host(c, &h, all, sizeof(h));
    h = 0;
    for (j = 0; j < c->size; j++) h += all[j];
    return h;
}

/* ------------------------------------------------------------------------- */
/* the distributed histogram sort on device buffers                          */

struct rank_info {
    int64_t nmemb, outnmemb;
    uint64_t kmin[MPS_MAX_KEY_WORDS], kmax[MPS_MAX_KEY_WORDS];
    uint64_t q01, q99;          /* single-word keys: the keys 1 % from either end of the sorted array */
};

/* gather everything on one leader, sort there, scatter by outnmemb
 * (reference: MPIU_Gather / MPIU_Scatter around the leaders' sort,
 *  mpsort-mpi.c:276-304; leader = rank with most data, lowest on ties, mp-mpiu.c:465) */
static void gather_sort(struct mpsort_comm * c, const void * dbase, void * dout, size_t elsize,
        const struct mpsort_radix_desc * desc, const int64_t * nmemb, const int64_t * outnmemb)
{
    const int p = c->size;
    int j, k, leader = 0;
    int64_t total = 0;
    for (j = 0; j < p; j++) {
        total += nmemb[j];
        if (nmemb[j] > nmemb[leader]) leader = j;
    }
    int64_t * cut = (int64_t *) malloc(sizeof(int64_t) * (size_t) p * (p + 1));
    mps_input_wait(c, dbase, 0, (size_t) nmemb[c->rank]);
    /* 1: everyone -> leader */
    for (j = 0; j < p; j++) for (k = 0; k <= p; k++) cut[(size_t) j * (p + 1) + k] = (k <= leader) ? 0 : nmemb[j];
    char * all = (char *) mps_arena_get(c, MPS_S_RECV, (size_t) (c->rank == leader ? total : 0) * elsize);
    uint64_t remote = 0, r2 = 0;
    mps_comm_alltoallv_dev(c, dbase, all, cut, elsize, 0, &remote);
    for (k = 0; k < p; k++) c->sendcounts[k] = (k == leader) ? nmemb[c->rank] : 0;
    timer_mark(c, "Gather");
    /* 2: leader sorts */
    char * sorted = (char *) mps_arena_get(c, MPS_S_SEND, (size_t) (c->rank == leader ? total : 0) * elsize);
    if (c->rank == leader) {
        struct sorted_view v;
        local_sort(c, all, (size_t) total, elsize, desc, 0, 1, &v);
        KERN_T(c, MPS_K_GATHER_RECORDS, mpsk_gather_records(all, v.idx, sorted, (size_t) total, elsize, c->stream));
        c->stats.first_sort_passes = v.npasses;
    }
    timer_mark(c, "FirstSort");
    /* 3: leader -> everyone by outnmemb */
    for (j = 0; j < p; j++) {
        int64_t acc = 0;
        for (k = 0; k <= p; k++) {
            cut[(size_t) j * (p + 1) + k] = (j == leader) ? acc : 0;
            if (k < p) acc += outnmemb[k];
        }
    }
    mps_comm_alltoallv_dev(c, sorted, dout, cut, elsize, 0, &r2);
    timer_mark(c, "Scatter");
    c->stats.bytes_sent_remote = remote + r2;
    c->stats.used_gather = 1;
    free(cut);
}

static void histogram_sort(struct mpsort_comm * c, const void * dbase, size_t n,
        void * dout, size_t outn, size_t elsize, const struct mpsort_radix_desc * desc)
{
    const int p = c->size;
    const uint32_t nw = key_words(desc);
    int j, b, w;
    struct sorted_view v1;

    /* ---- sizes first: needed for the gather decision and the abort on mismatch.
     * (MPIU_Segmenter_collect_sizes x2, mp-mpiu.c:394-425) */
    struct rank_info mine, * info = (struct rank_info *) malloc(sizeof(struct rank_info) * (size_t) p);
    memset(&mine, 0, sizeof(mine));
    mine.nmemb = (int64_t) n;
    mine.outnmemb = (int64_t) outn;
    int64_t nmemb[MPS_MAX_RANKS], outnmemb[MPS_MAX_RANKS], C[MPS_MAX_RANKS + 1];
    int64_t total = 0, totalout = 0;
    if (p > 1) {
        int64_t sz[2] = { (int64_t) n, (int64_t) outn }, allsz[2 * MPS_MAX_RANKS];
        mpsort_comm_allgather_host(c, sz, allsz, sizeof(sz));
        for (j = 0; j < p; j++) { nmemb[j] = allsz[2 * j]; outnmemb[j] = allsz[2 * j + 1]; }
    } else {
        nmemb[0] = (int64_t) n; outnmemb[0] = (int64_t) outn;
    }
    for (j = 0; j < p; j++) { total += nmemb[j]; totalout += outnmemb[j]; }
    if (total != totalout) {
        /* mpsort-mpi.c:225-232 */
        if (c->rank == 0)
            fprintf(stderr, "Input and output size mismatch: %td (in) != %td (out)"
                            "Caller site: %s:%d\n", (ptrdiff_t) total, (ptrdiff_t) totalout,
                            mps_caller_file, mps_caller_line);
        mps_fatal(c, __FILE__, __LINE__, "total number of items in the output does not match the input");
    }
    mpsort_cumulative_counts(p, outnmemb, C);

    timer_mark(c, "START");

    /* ---- small-input path (mpsort-mpi.c:234-256): the reference merges tiny ranks
     * into segments sorted by one leader; on GPUs the whole job is latency bound
     * below a few thousand records, so everything goes to one leader. */
    if (p > 1) {
        int use_gather = 0;
        if (mpsort_mpi_has_options(MPSORT_REQUIRE_GATHER_SORT)) {
            use_gather = 1;
            if (c->rank == 0)
                fprintf(stderr, "MPSort: gathering all data to a single rank for sorting due to MPSORT_REQUIRE_GATHER_SORT. "
                                "Total number of items is %ld. Caller site: %s:%d\n",
                                (long) total, mps_caller_file, mps_caller_line);
        } else if (mpsort_mpi_has_options(MPSORT_DISABLE_GATHER_SORT)) {
            use_gather = 0;
            if (c->rank == 0)
                fprintf(stderr, "MPSort: disable gathering data into larger chunks due to MPSORT_DISABLE_GATHER_SORT. "
                                "Caller site: %s:%d\n", mps_caller_file, mps_caller_line);
        } else {
            /* the reference's own cap: no more than 4 MiB in a segment (:235-238) */
            use_gather = (total <= 65536) && ((size_t) total * elsize <= ((size_t) 4 << 20));
        }
        if (use_gather) {
            gather_sort(c, dbase, dout, elsize, desc, nmemb, outnmemb);
            timer_mark(c, "END");
            free(info);
            return;
        }
    }

    /* ---- FirstSort */
    void * sendbuf = NULL;
    if (p == 1 && rec16_applicable(elsize, desc, dbase, dout)) {
        /* one rank, record mode: the passes leave the sorted records in the output */
        local_sort_rec16(c, dbase, n, elsize, desc, dout, &v1);
        c->stats.first_sort_passes = v1.npasses;
        c->stats.record_mode = 1;
        c->sendcounts[0] = (int64_t) n;
        timer_mark(c, "FirstSort");
        timer_mark(c, "END");
        free(info);
        return;
    }
    /* record mode sorts straight into the send buffer; index mode only needs one when it packs (no fused
     * pack) or when peers pull from it, so it is made then */
    if (p > 1 && (rec16_shape(elsize, desc) || c->p2p.pull || c->kind != MPS_T_NCCL || c->p2p.disabled))
        sendbuf = mps_arena_get(c, MPS_S_SEND, n * elsize);
    if (p > 1 && sendbuf && rec16_applicable(elsize, desc, dbase, sendbuf)) {
        local_sort_rec16(c, dbase, n, elsize, desc, sendbuf, &v1);
        c->stats.record_mode = 1;
    } else {
        local_sort(c, dbase, n, elsize, desc, p > 1, 1, &v1);
    }
    c->stats.first_sort_passes = v1.npasses;
    c->stats.key_words = nw;

    if (p == 1) {
        /* one rank: the exchange is a self copy and the second sort the identity
         * (SURVEY.md appendix B.6): gather straight into the output */
        if (dout != dbase) {
            KERN_T(c, MPS_K_GATHER_RECORDS, mpsk_gather_records(dbase, v1.idx, dout, n, elsize, c->stream));
        } else {
            void * tmp = mps_arena_get(c, MPS_S_SEND, n * elsize);
            KERN_T(c, MPS_K_GATHER_RECORDS, mpsk_gather_records(dbase, v1.idx, tmp, n, elsize, c->stream));
            if (n) CUDA_OK(c, cudaMemcpyAsync(dout, tmp, n * elsize, cudaMemcpyDeviceToDevice, c->stream));
        }
        c->sendcounts[0] = (int64_t) n;
        timer_mark(c, "FirstSort");
        timer_mark(c, "END");
        free(info);
        return;
    }
    timer_mark(c, "FirstSort");

    /* ---- PmaxPmin (mpsort-mpi.c:606-661): ends of the locally sorted keys */
    if (n > 0) {
        const char * kb = (const char *) v1.kv.base;
        uint64_t * h = (uint64_t *) mps_host_stage(c, (2 * MPS_MAX_KEY_WORDS + 2) * sizeof(uint64_t));
        for (w = 0; w < (int) nw; w++) {
            CUDA_OK(c, cudaMemcpyAsync(h + w, kb + (size_t) w * v1.kv.word_stride, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
            CUDA_OK(c, cudaMemcpyAsync(h + MPS_MAX_KEY_WORDS + w, kb + (n - 1) * v1.kv.item_stride + (size_t) w * v1.kv.word_stride,
                                       sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        }
        /* (two more keys for the sparse-exchange hint below) */
        CUDA_OK(c, cudaMemcpyAsync(h + 2 * MPS_MAX_KEY_WORDS, kb + (n / 100) * v1.kv.item_stride, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(c, cudaMemcpyAsync(h + 2 * MPS_MAX_KEY_WORDS + 1, kb + (n - 1 - n / 100) * v1.kv.item_stride, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(c, cudaStreamSynchronize(c->stream));
        for (w = 0; w < (int) nw; w++) {
            mine.kmin[w] = (h[w] ^ v1.kv.flip) + (w == 0 ? v1.kv.add : 0);
            mine.kmax[w] = (h[MPS_MAX_KEY_WORDS + w] ^ v1.kv.flip) + (w == 0 ? v1.kv.add : 0);
        }
        mine.q01 = (h[2 * MPS_MAX_KEY_WORDS] ^ v1.kv.flip) + v1.kv.add;
        mine.q99 = (h[2 * MPS_MAX_KEY_WORDS + 1] ^ v1.kv.flip) + v1.kv.add;
    }
    mpsort_comm_allgather_host(c, &mine, info, sizeof(mine));
    uint64_t kmin[MPS_MAX_RANKS * MPS_MAX_KEY_WORDS], kmax[MPS_MAX_RANKS * MPS_MAX_KEY_WORDS];
    uint64_t Pmin[MPS_MAX_KEY_WORDS], Pmax[MPS_MAX_KEY_WORDS], prefix0[MPS_MAX_KEY_WORDS];
    for (j = 0; j < p; j++) for (w = 0; w < (int) nw; w++) {
        kmin[(size_t) j * nw + w] = info[j].kmin[w];
        kmax[(size_t) j * nw + w] = info[j].kmax[w];
    }
    const int level0 = mpsort_key_range(p, nw, nmemb, kmin, kmax, Pmin, Pmax, prefix0);
    timer_mark(c, "PmaxPmin");

    /* ---- pipelining (MPSORT_EXCHANGE_PHASES=Q): every rank's output is cut into Q
     * consecutive parts ("virtual ranks", pv = p*Q destinations). The exchange then runs
     * part by part and the merge of part q overlaps the transfer of part q+1 on a second
     * stream. The result is the same global stable sort: only more cut points. Q is
     * decided from global facts so that every rank takes the same branch. */
    int Q = 1;
    {
        /* default: several parts when the slices move by DMA or by the fused pack kernel (the merge of a part
         * overlaps the transfer of the next, and only the last part's merge is exposed), one when grouped
         * send/recv does the moving (transfer and merge then slow each other down, profiles/r01_pipelined_exchange.log).
         * Measured with 2^28 16-byte records per GPU (profiles/r02_call_n2_*.log, r02_call_n8_*.log), ms per sort for
         * 1 / 2 / 3 / 4 / 8 parts: 8 GPUs 19.6 / 17.4 / 16.7 / 16.4 / 16.3; 2 GPUs 15.2 / 14.2 / - / 13.8 / 13.9;
         * 4 GPUs (final kernels, profiles/r02_call_n4_parts.log) 4 / 6 / 8 parts: 14.52 / 14.36 / 14.34.
         * 48-byte records at 8 GPUs: 43.5 with 2 parts, 40.2 with 4. */
        const char * e = getenv("MPSORT_EXCHANGE_PHASES");
        const int dma = c->kind == MPS_T_NCCL && !c->p2p.disabled && c->p2p.copy_engine > 0 && !c->p2p.pull;
        /* A hint that the exchange will be SPARSE (mostly sorted input): every rank keeps its record count and the
         * ranks' key ranges, 1 % trimmed at either end, do not overlap and follow rank order. Then almost nothing
         * moves, there is no transfer for the merges to hide behind, and parts only cost launches: two are taken
         * (16.5 ms with 8 parts, 16.1 with 2 or 1 for config 4 at 8 GPUs, profiles/r02_call_n8_*.log). The hint only
         * picks the number of parts; a wrong one costs time, never the result. */
        int sparse_hint = nw == 1;
        {
            int prev = -1;
            for (j = 0; j < p && sparse_hint; j++) {
                if (nmemb[j] != outnmemb[j]) sparse_hint = 0;
                if (nmemb[j] == 0) continue;
                if (prev >= 0 && info[prev].q99 > info[j].q01) sparse_hint = 0;
                prev = j;
            }
        }
        const int want = e ? atoi(e) : (dma ? (sparse_hint ? 2 : ((p >= 4 && (elsize == 16 || elsize == 8)) ? 8 : 4)) : 1);
        /* parts only pay for themselves on large inputs; MPSORT_PHASES_MIN_RECORDS (records per rank,
         * default 2^22) moves the threshold -- the CPU host-flow tests use it to cut tiny inputs */
        const char * m = getenv("MPSORT_PHASES_MIN_RECORDS");
        const int64_t min_per_rank = m ? (int64_t) atoll(m) : ((int64_t) 1 << 22);
        if (nw == 1 && p <= 32 && want > 1 && total / p >= min_per_rank) {
            Q = want;
            while (Q > 1 && p * Q > MPS_MAX_RANKS) Q--;
        }
    }
    const int pv = p * Q;
    int64_t Cv[MPS_MAX_RANKS + 1];
#define PARTBASE(k, q) ((int64_t) (outnmemb[k] * (int64_t) (q) / Q))
    Cv[0] = 0;
    for (j = 0; j < p; j++) for (b = 0; b < Q; b++)
        Cv[j * Q + b + 1] = Cv[j * Q + b] + (PARTBASE(j, b + 1) - PARTBASE(j, b));

    /* ---- findP: byte-wise descent to the key at global rank Cv[b]-1 for every
     * boundary b. Any exact selection gives the reference's result because a
     * splitter is only accepted when CLT < C <= CLE (internal-parallel.h:234-247),
     * i.e. when it IS that key. */
    const int ns = pv - 1;
    const int nlevels = 8 * (int) nw;
    /* layout of the splitter slot: prefix[ns][nw] | target[ns] | counts[ns][256] | final[2*ns] */
    const size_t sp_words = (size_t) ns * nw + ns + (size_t) ns * 256 + 2 * (size_t) ns;
    uint64_t * d_sp = (uint64_t *) mps_arena_get(c, MPS_S_SPLIT, sp_words * sizeof(uint64_t));
    uint64_t * d_prefix = d_sp;
    uint64_t * d_target = d_prefix + (size_t) ns * nw;
    uint64_t * d_counts = d_target + ns;
    uint64_t * d_final = d_counts + (size_t) ns * 256;
    {
        uint64_t * h = (uint64_t *) mps_host_stage(c, ((size_t) ns * nw + ns) * sizeof(uint64_t));
        for (b = 0; b < ns; b++) {
            for (w = 0; w < (int) nw; w++) h[(size_t) b * nw + w] = prefix0[w];
            h[(size_t) ns * nw + b] = (uint64_t) Cv[b + 1];
        }
        CUDA_OK(c, cudaMemcpyAsync(d_prefix, h, ((size_t) ns * nw + ns) * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    }
    int level, round = 0;
    /* One process per GPU (NCCL transport): every level in ONE kernel per GPU, the per-level sums taken over
     * mapped peer memory (mpsk_splitter_descent_peer) instead of count kernel + ncclAllReduce + select kernel
     * per level: 17.42 -> 17.16 ms per sort at 8 GPUs, 14.15 -> 13.92 at 2 (profiles/r02_call_n8_*.log,
     * r02_call_n2_*.log). MPSORT_NO_PEER_SPLITTER=1 keeps the all-reduce per level, which is also what rank
     * threads of one process use (MPSORT_PEER_SPLITTER=1 turns the kernel on for them too) and the fallback
     * when a mailbox cannot be mapped -- a collective verdict. The switches must agree on all ranks. */
    const int peer_wanted = c->kind == MPS_T_NCCL ? getenv("MPSORT_NO_PEER_SPLITTER") == NULL : getenv("MPSORT_PEER_SPLITTER") != NULL;
    const int peer_desc = ns > 0 && ns <= 63 && level0 < nlevels && peer_wanted && mps_comm_peer_boxes_prepare(c);
    if (peer_desc) {
        KERN_T(c, MPS_K_SPLITTER, mps_comm_peer_descent(c, v1.kv, n, nw, d_prefix, d_target, ns, level0, nlevels));
        mps_comm_peer_descent_check(c);
        round = nlevels - level0;
        timer_mark(c, "bisect0001");
    }
    for (level = level0; level < nlevels && !peer_desc; level++) {
        KERN_T(c, MPS_K_SPLITTER, mpsk_splitter_count(v1.kv, n, nw, d_prefix, ns, level, d_counts, c->stream));
        mps_comm_allreduce_u64_dev(c, d_counts, (size_t) ns * 256);
        KERN_T(c, MPS_K_SPLITTER, mpsk_splitter_select(d_counts, d_target, d_prefix, nw, ns, level, c->stream));
        round++;
        if (round <= 10) {
            char name[20];
            snprintf(name, sizeof(name), "bisect%04d", round);
            timer_mark(c, name);
        }
    }
    c->stats.splitter_rounds = (uint32_t) round;
    KERN_T(c, MPS_K_SPLITTER, mpsk_splitter_final(v1.kv, n, nw, d_prefix, ns, d_final, c->stream));
    timer_mark(c, "findP");

    /* ---- LayDistr: all-gather the local rows (replaces the 8-byte Alltoalls :450-456) */
    int64_t * rows = (int64_t *) malloc(sizeof(int64_t) * 2 * (size_t) ns * p);
    void * recvbuf = mps_arena_get(c, MPS_S_RECV, outn * elsize);
    int use_p2p = 0;
    {
        /* one message per rank: its (CLT, CLE) row and how to reach its receive buffer */
        struct lay_msg { int64_t row[2 * MPS_MAX_RANKS]; struct mps_recv_info recv; };
        struct lay_msg mymsg, * allmsg = (struct lay_msg *) malloc(sizeof(struct lay_msg) * (size_t) p);
        struct mps_recv_info * allrecv = (struct mps_recv_info *) malloc(sizeof(struct mps_recv_info) * (size_t) p);
        int64_t * h = (int64_t *) mps_host_stage(c, 2 * (size_t) ns * sizeof(int64_t));
        CUDA_OK(c, cudaMemcpyAsync(h, d_final, 2 * (size_t) ns * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(c, cudaStreamSynchronize(c->stream));
        memset(&mymsg, 0, sizeof(mymsg));
        memcpy(mymsg.row, h, 2 * (size_t) ns * sizeof(int64_t));
        mps_comm_recv_info(c, recvbuf, sendbuf, &mymsg.recv);
        mpsort_comm_allgather_host(c, &mymsg, allmsg, sizeof(mymsg));
        for (j = 0; j < p; j++) {
            memcpy(rows + (size_t) j * 2 * ns, allmsg[j].row, 2 * (size_t) ns * sizeof(int64_t));
            allrecv[j] = allmsg[j].recv;
        }
        use_p2p = mps_comm_p2p_prepare(c, allrecv);
        free(allmsg); free(allrecv);
    }
    timer_mark(c, "LayDistr");

    /* ---- LaySolve: p sources x pv virtual destinations */
    int64_t * clt = (int64_t *) malloc(sizeof(int64_t) * (size_t) (ns ? ns : 1) * p);
    int64_t * cle = (int64_t *) malloc(sizeof(int64_t) * (size_t) (ns ? ns : 1) * p);
    int64_t * cut = (int64_t *) malloc(sizeof(int64_t) * (size_t) p * (pv + 1));
    for (j = 0; j < p; j++) for (b = 0; b < ns; b++) {
        clt[(size_t) j * ns + b] = rows[(size_t) j * 2 * ns + b];
        cle[(size_t) j * ns + b] = rows[(size_t) j * 2 * ns + ns + b];
    }
    {
        const int rc = mpsort_solve_layout2(p, pv, Cv, clt, cle, nmemb, cut);
        if (rc != 0) mps_fatal(c, __FILE__, __LINE__, "serious bug: layout solver failed with code %d", rc);
    }
#define CUTV(j, v) cut[(size_t) (j) * (pv + 1) + (v)]
    /* consistency checks of mpsort-mpi.c:490-510 */
    {
        int64_t totrecv = 0;
        for (j = 0; j < p; j++) totrecv += CUTV(j, (c->rank + 1) * Q) - CUTV(j, c->rank * Q);
        if (totrecv != (int64_t) outn)
            mps_fatal(c, __FILE__, __LINE__, "totrecv = %td, mismatch with %td", (ptrdiff_t) totrecv, (ptrdiff_t) outn);
    }
    /* the reference's SendCount row (mpsort-mpi.c:483-485): all parts of a rank together */
    for (j = 0; j < p; j++) c->sendcounts[j] = CUTV(c->rank, (j + 1) * Q) - CUTV(c->rank, j * Q);
    timer_mark(c, "LaySolve");

    /* ---- Exchange: pack (payload gather into destination-contiguous order; the
     * destinations are contiguous slices of the sorted order, SendDispl[i] == myC[i]
     * :483-501) then, part by part, grouped send/recv or peer stores */
    /* Index mode over mapped peer buffers skips the pack: ONE kernel gathers by sorted index and stores into
     * the peers' receive buffers over NVLink (mps_comm_exchange_gather) -- no send buffer, no separate 2E+4
     * bytes-per-record pack pass. 48-byte particles at 8 GPUs, 4 parts: 40.2 ms with pack + DMA copies, 36.2 with
     * the pack of part q+1 beside the copy of part q, 35.7 fused (profiles/r02_call_n8_*.log); at 2 GPUs 29.7 /
     * 27.5 / 26.3. MPSORT_NO_FUSED_PACK=1 packs first. use_p2p is a collective verdict, so is this. */
    const int index_mode = v1.sorted_recs == NULL;
    const int fused_pack = index_mode && use_p2p && !c->p2p.pull && n <= 0xffffffffu && getenv("MPSORT_NO_FUSED_PACK") == NULL;
    if (index_mode && !fused_pack) {
        if (!sendbuf) sendbuf = mps_arena_get(c, MPS_S_SEND, n * elsize);
        KERN_T(c, MPS_K_GATHER_RECORDS, mpsk_gather_records(dbase, v1.idx, sendbuf, n, elsize, c->stream));
    }
    timer_mark(c, "Pack");
    mps_merge_ovf_begin(c);
    const int dense = mpsort_mpi_has_options(MPSORT_DISABLE_SPARSE_ALLTOALLV)
                      && !mpsort_mpi_has_options(MPSORT_REQUIRE_SPARSE_ALLTOALLV);
    c->stats.dense_exchange = (uint32_t) dense;
    c->stats.p2p_exchange = (uint32_t) use_p2p;
    c->stats.exchange_phases = (uint32_t) Q;
    if (!c->phase_ev_created) {
        for (j = 0; j <= MPS_MAX_RANKS; j++) CUDA_OK(c, cudaEventCreateWithFlags(&c->phase_ev[j], cudaEventDisableTiming));
        c->phase_ev_created = 1;
    }
    if (Q > 1 && !c->stream2) CUDA_OK(c, cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    const int me = c->rank;
    int64_t (*recvcnt_q)[MPS_MAX_RANKS] = (int64_t (*)[MPS_MAX_RANKS]) malloc(sizeof(int64_t) * MPS_MAX_RANKS * (size_t) Q);
    int self_direct[MPS_MAX_RANKS];
    const char * self_ptr[MPS_MAX_RANKS];
    int q;
    /* A SPARSE exchange (mostly sorted input: 21 MB leave a GPU of the 4 GiB it holds) is all latency: part after
     * part, each with its own completion barrier, took 1.4 ms at 8 GPUs with two parts. When no rank sends more than
     * 128 MiB away, all parts move in ONE step -- every copy queued at once, one barrier -- and the merges follow.
     * Every rank holds the whole cut matrix, so all take the same branch. */
    /* how the copies of part q+1 follow those of part q (exchange_p2p): back to back with one peer, in lock
     * step behind every part's barrier with more (measured); MPSORT_CHAINED_PARTS=0/1/2 overrides */
    const char * chain_env = getenv("MPSORT_CHAINED_PARTS");
    const int chain_mode = getenv("MPSORT_NO_CHAINED_PARTS") ? 0 : (chain_env ? atoi(chain_env) : (p == 2 ? 1 : 0));
    int one_step = 0;
    if (Q > 1 && use_p2p) {
        int64_t worst = 0;
        for (j = 0; j < p; j++) {
            const int64_t away = nmemb[j] - (CUTV(j, (j + 1) * Q) - CUTV(j, j * Q));
            if (away > worst) worst = away;
        }
        one_step = (uint64_t) worst * elsize <= ((uint64_t) 128 << 20) && !getenv("MPSORT_NO_ONE_STEP");
    }
    for (q = 0; q < Q; q++) {
        int64_t sendoff[MPS_MAX_RANKS], sendcnt[MPS_MAX_RANKS], peer_recvoff[MPS_MAX_RANKS], peer_sendoff[MPS_MAX_RANKS];
        int k;
        for (k = 0; k < p; k++) {
            const int v = k * Q + q;                   /* part q of rank k */
            sendoff[k] = CUTV(me, v);
            sendcnt[k] = CUTV(me, v + 1) - CUTV(me, v);
            peer_recvoff[k] = PARTBASE(k, q);
            for (j = 0; j < me; j++) peer_recvoff[k] += CUTV(j, v + 1) - CUTV(j, v);
            recvcnt_q[q][k] = CUTV(k, me * Q + q + 1) - CUTV(k, me * Q + q);
            peer_sendoff[k] = CUTV(k, me * Q + q);
        }
        /* My own slice of a part that will be MERGED stays where it is: the merge reads run `me` from the send
         * buffer (sorted records, or the packed ones), no transport copies it. Mostly sorted input keeps 99 % of
         * its records: that copy alone was 1.3 ms of 14.7 (profiles/r02_call_n8_c_*.log). A part that takes the
         * radix fallback needs its runs side by side and gets the copy. */
        {
            const int64_t cnt_q = PARTBASE(me, q + 1) - PARTBASE(me, q);
            self_direct[q] = sendbuf != NULL && !fused_pack && recvcnt_q[q][me] > 0 && cnt_q < ((int64_t) 1 << 31)
                             && merge_applicable(p, (size_t) cnt_q, desc) && !getenv("MPSORT_NO_SELF_IN_PLACE");
            self_ptr[q] = self_direct[q] ? (const char *) sendbuf + (size_t) sendoff[me] * elsize : NULL;
            c->stats.own_slices_in_place += (uint32_t) self_direct[q];
        }
        c->p2p.skip_self = self_direct[q];
        c->p2p.skip_barrier = one_step && q + 1 < Q;
        c->p2p.burst = one_step;            /* many small copies: on all copy streams at once */
        c->p2p.part = q;
        c->p2p.chained = q > 0 ? chain_mode : 0;
        mps_kt_begin(c, MPS_K_EXCHANGE);
        if (fused_pack)
            mps_comm_exchange_gather(c, dbase, v1.idx, sendoff, sendcnt, recvbuf, peer_recvoff, elsize, &c->stats.bytes_sent_remote);
        else
            mps_comm_exchange(c, sendbuf, sendoff, sendcnt, recvbuf, PARTBASE(me, q), recvcnt_q[q], peer_recvoff, peer_sendoff,
                              elsize, dense, use_p2p, &c->stats.bytes_sent_remote);
        mps_kt_end(c);
        c->p2p.skip_barrier = 0;
        c->p2p.burst = 0;
        c->p2p.chained = 0;
        c->p2p.part = 0;
        c->p2p.skip_self = 0;
        /* (one step: no part is complete before the barrier that follows the last) */
        if (!one_step) CUDA_OK(c, cudaEventRecord(c->phase_ev[q], c->stream));
    }
    if (one_step) for (q = 0; q < Q; q++) CUDA_OK(c, cudaEventRecord(c->phase_ev[q], c->stream));
    timer_mark(c, "Exchange");

    /* ---- SecondSort: every received part is p sorted runs in source-rank order; a stable
     * merge (or stable re-sort) of it restores global order with ties by (source rank,
     * index). With Q > 1 the parts are merged on the second stream as they arrive. */
    {
        cudaStream_t main_stream = c->stream;
        if (Q > 1) c->stream = c->stream2;
        for (q = 0; q < Q; q++) {
            const int64_t base = PARTBASE(me, q), cnt = PARTBASE(me, q + 1) - PARTBASE(me, q);
            int64_t rdispl[MPS_MAX_RANKS + 1];
            rdispl[0] = 0;
            for (j = 0; j < p; j++) rdispl[j + 1] = rdispl[j] + recvcnt_q[q][j];
            /* (the fused pack reads the input and the sorted permutation until its last part is out: an
             * in-place sort, whose merges write that very input, and a part that is re-sorted by the radix
             * path, which reuses the arena slots the permutation lives in, must not start before that) */
            const int late = fused_pack && (dout == dbase || !merge_applicable(p, (size_t) cnt, desc));
            if (Q > 1) CUDA_OK(c, cudaStreamWaitEvent(c->stream, c->phase_ev[late ? Q - 1 : q], 0));
            char * part_in = (char *) recvbuf + (size_t) base * elsize;
            char * part_out = (char *) dout + (size_t) base * elsize;
            if (merge_received_runs(c, p, part_in, rdispl, part_out, (size_t) cnt, elsize, desc,
                                    self_direct[q] ? me : -1, self_ptr[q]) != 0) {
                struct sorted_view v2;
                if (self_direct[q])          /* (cannot happen: the same test chose both; never leave a hole) */
                    CUDA_OK(c, cudaMemcpyAsync(part_in + (size_t) rdispl[me] * elsize, self_ptr[q], (size_t) recvcnt_q[q][me] * elsize,
                                               cudaMemcpyDeviceToDevice, c->stream));
                local_sort(c, part_in, (size_t) cnt, elsize, desc, 0, 1, &v2);
                KERN_T(c, MPS_K_GATHER_RECORDS, mpsk_gather_records(part_in, v2.idx, part_out, (size_t) cnt, elsize, c->stream));
                c->stats.second_sort_passes = v2.npasses;
            }
            /* a host output leaves part by part, beside the merge of the next part */
            if (q + 1 < Q) mps_output_ready(c, (size_t) (base + cnt));
        }
        if (Q > 1) {
            CUDA_OK(c, cudaEventRecord(c->phase_ev[MPS_MAX_RANKS], c->stream));
            c->stream = main_stream;
            CUDA_OK(c, cudaStreamWaitEvent(c->stream, c->phase_ev[MPS_MAX_RANKS], 0));
        }
    }
    mps_merge_ovf_fetch(c);
    timer_mark(c, "SecondSort");
    timer_mark(c, "END");
#undef CUTV
#undef PARTBASE

    free(recvcnt_q);
    free(rows); free(clt); free(cle); free(cut); free(info);
}

/* ------------------------------------------------------------------------- */
/* public entry points                                                        */

static void validate(struct mpsort_comm * c, size_t elsize, const struct mpsort_radix_desc * d)
{
    if (!c) { fprintf(stderr, "MPSort: NULL communicator. Caller site: %s:%d\n", mps_caller_file, mps_caller_line); abort(); }
    if (!d) mps_fatal(c, __FILE__, __LINE__, "NULL radix descriptor");
    if (!(d->width == 1 || d->width == 2 || d->width == 4 || d->width == 8))
        mps_fatal(c, __FILE__, __LINE__, "radix word width %u is not 1, 2, 4 or 8", d->width);
    if (d->nwords < 1) mps_fatal(c, __FILE__, __LINE__, "radix descriptor has no words");
    if (d->offset + (size_t) d->width * d->nwords > elsize)
        mps_fatal(c, __FILE__, __LINE__, "radix key [%zu, %zu) does not fit in a %zu byte element",
                  d->offset, d->offset + (size_t) d->width * d->nwords, elsize);
    if (key_words(d) > MPS_MAX_KEY_WORDS)
        mps_fatal(c, __FILE__, __LINE__, "radix size %zu bytes exceeds the supported maximum %d",
                  (size_t) d->width * d->nwords, MPS_MAX_KEY_WORDS * 8);
    if (elsize == 0) mps_fatal(c, __FILE__, __LINE__, "element size is zero");
}

void mpsort_mpi_newarray_desc_impl(void * base, size_t nmemb,
        void * out, size_t outnmemb, size_t elsize,
        const struct mpsort_radix_desc * desc, mpsort_comm_t c,
        const int line, const char * file)
{
    mps_caller_file = file ? file : "?";
    mps_caller_line = line;
    validate(c, elsize, desc);
    CUDA_OK(c, cudaSetDevice(c->device));

    memset(&c->stats, 0, sizeof(c->stats));
    memset(c->sendcounts, 0, sizeof(c->sendcounts));
    c->stats.nmemb = nmemb; c->stats.outnmemb = outnmemb; c->stats.elsize = elsize;
    c->stats.key_words = key_words(desc);
    timer_reset(c);

    /* host or device buffers? */
    const int in_dev = (nmemb == 0) ? 1 : is_device_pointer(c, base);
    const int out_dev = (outnmemb == 0) ? 1 : is_device_pointer(c, out);
    const void * dbase = base;
    void * dout = out;
    c->in.dbase = NULL;
    c->outp.host = NULL;
    if (!in_dev) {
        void * din = mps_arena_get(c, MPS_S_DIN, nmemb * elsize);
        input_stage(c, din, base, nmemb, elsize);
        dbase = din;
    }
    if (!out_dev) {
        /* a separate staged output also for host in-place: saves the copy-back pass */
        dout = mps_arena_get(c, MPS_S_DOUT, outnmemb * elsize);
    } else if (out == base) {
        dout = (void *) dbase;
    }

    uint64_t sum1 = 0;
    const int verify = mpsort_mpi_has_options(MPSORT_VERIFY_CHECKSUM);
    if (verify) {
        mps_input_wait(c, dbase, 0, nmemb);
        sum1 = device_checksum(c, dbase, nmemb * elsize);
    }
    /* (with the checksum on, the output is only sent once it has been verified) */
    if (!out_dev && !verify) output_begin(c, out, dout, outnmemb, elsize);

    histogram_sort(c, dbase, nmemb, dout, outnmemb, elsize, desc);
    mps_input_wait(c, dbase, 0, nmemb);            /* (paths that never looked at the input: empty ranks) */
    c->in.dbase = NULL;

    if (verify) {
        const uint64_t sum2 = device_checksum(c, dout, outnmemb * elsize);
        if (sum1 != sum2)
            mps_fatal(c, __FILE__, __LINE__, "Data changed after sorting; checksum mismatch");   /* :324-330 */
    }
    if (c->outp.host) output_finish(c);
    else if (!out_dev && outnmemb > 0)
        CUDA_OK(c, cudaMemcpyAsync(out, dout, outnmemb * elsize, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    mps_merge_ovf_check(c);
    mps_kt_collect(c);
    timer_publish(c);
}

void mpsort_mpi_desc_impl(void * base, size_t nmemb, size_t elsize,
        const struct mpsort_radix_desc * desc, mpsort_comm_t comm,
        const int line, const char * file)
{
    mpsort_mpi_newarray_desc_impl(base, nmemb, base, nmemb, elsize, desc, comm, line, file);
}

/* one cached size-1 communicator (stream + arena) per device and thread, shared by
 * radix_sort_desc and radix_sort. Its arenas only grow (several times nmemb * size after a large
 * call): mpsort_release_cached() gives them back for the calling thread; a thread that exits
 * releases its own. */
static pthread_key_t g_cache_key;
static pthread_once_t g_cache_once = PTHREAD_ONCE_INIT;

static void cache_destroy(void * p)
{
    mpsort_comm_t * self = (mpsort_comm_t *) p;
    int d;
    if (!self) return;
    for (d = 0; d < 64; d++) if (self[d]) mpsort_comm_destroy(self[d]);
    free(self);
}

static void cache_make_key(void) { pthread_key_create(&g_cache_key, cache_destroy); }

static mpsort_comm_t cached_self_comm(int device)
{
    if (device < 0 || device >= 64) { fprintf(stderr, "MPSort: bad device %d\n", device); abort(); }
    pthread_once(&g_cache_once, cache_make_key);
    mpsort_comm_t * self = (mpsort_comm_t *) pthread_getspecific(g_cache_key);
    if (!self) {
        self = (mpsort_comm_t *) calloc(64, sizeof(mpsort_comm_t));
        if (!self) { fprintf(stderr, "MPSort: out of host memory\n"); abort(); }
        pthread_setspecific(g_cache_key, self);
    }
    if (!self[device]) self[device] = mpsort_comm_self(device);
    return self[device];
}

void mpsort_release_cached(void)
{
    pthread_once(&g_cache_once, cache_make_key);
    void * p = pthread_getspecific(g_cache_key);
    if (p) { pthread_setspecific(g_cache_key, NULL); cache_destroy(p); }
}

void radix_sort_desc(void * base, size_t nmemb, size_t size,
        const struct mpsort_radix_desc * desc, int device)
{
    mpsort_mpi_newarray_desc_impl(base, nmemb, base, nmemb, size, desc, cached_self_comm(device), __LINE__, __FILE__);
}

/* ------------------------------------------------------------------------- */
/* the reference's own signatures: host radix() callback + rsize + arg        */
/*
 * mpsort_mpi_impl / mpsort_mpi_newarray_impl / radix_sort exactly as declared in the
 * reference's mpsort.h:1-4,25-47, for callers whose key is not one contiguous field
 * (test-issue7.c:12-17 builds it from two) and that cannot switch to a descriptor.
 * A host function pointer cannot run on the GPU, so the callback is evaluated once per
 * record ON THE HOST (the reference evaluates it twice per comparison,
 * radixsort.c:19-26) into the front of an augmented record {radix | record}; the
 * augmented records are sorted on the GPU by the descriptor of that radix and the
 * records are copied back out. The radix orders as a little-endian integer of rsize
 * bytes (radixsort.c:47-98,178-193 on a little-endian host): rsize 2/4/8 are native
 * words, a multiple of 8 is u64 words with the last one most significant, anything
 * else compares byte-wise from the last byte down.
 * Buffers must be host memory here; device-resident data needs the descriptor API.
 * The reference's warnings about element / radix sizes that are large and not a
 * multiple of 8 (mpsort-mpi.c:200-215) concern MPI datatypes and are not repeated.
 */
int mpsort_callback_desc(size_t rsize, struct mpsort_radix_desc * d, size_t * rpad)
{
    if (rsize == 0 || rsize > (size_t) MPS_MAX_KEY_WORDS * 8) return -1;
    memset(d, 0, sizeof(*d));
    d->offset = 0;
    if (rsize % 8 == 0) { d->width = 8; d->nwords = (uint32_t) (rsize / 8); }
    else if (rsize == 4 || rsize == 2) { d->width = (uint32_t) rsize; d->nwords = 1; }
    else { d->width = 1; d->nwords = (uint32_t) rsize; }
    *rpad = (rsize + 7) & ~(size_t) 7;       /* keeps the record part 8-byte aligned */
    return 0;
}

void mpsort_callback_pack(const void * base, size_t nmemb, size_t elsize,
        mpsort_radix_func radix, size_t rsize, void * arg, void * aug)
{
    const size_t rpad = (rsize + 7) & ~(size_t) 7, E2 = rpad + elsize;
    const char * src = (const char *) base;
    char * dst = (char *) aug;
    size_t i;
    for (i = 0; i < nmemb; i++) {
        if (rpad != rsize) memset(dst + rsize, 0, rpad - rsize);
        radix(src, dst, arg);
        memcpy(dst + rpad, src, elsize);
        src += elsize;
        dst += E2;
    }
}

void mpsort_callback_unpack(const void * aug, size_t nmemb, size_t elsize, size_t rsize, void * out)
{
    const size_t rpad = (rsize + 7) & ~(size_t) 7, E2 = rpad + elsize;
    const char * src = (const char *) aug + rpad;
    char * dst = (char *) out;
    size_t i;
    for (i = 0; i < nmemb; i++) {
        memcpy(dst, src, elsize);
        src += E2;
        dst += elsize;
    }
}

void mpsort_mpi_newarray_impl(void * base, size_t nmemb,
        void * out, size_t outnmemb, size_t elsize,
        mpsort_radix_func radix, size_t rsize, void * arg,
        mpsort_comm_t c, const int line, const char * file)
{
    struct mpsort_radix_desc d;
    size_t rpad = 0;
    mps_caller_file = file ? file : "?";
    mps_caller_line = line;
    if (!c) { fprintf(stderr, "MPSort: NULL communicator. Caller site: %s:%d\n", mps_caller_file, mps_caller_line); abort(); }
    if (!radix) mps_fatal(c, __FILE__, __LINE__, "NULL radix callback");
    if (elsize == 0) mps_fatal(c, __FILE__, __LINE__, "element size is zero");
    if (mpsort_callback_desc(rsize, &d, &rpad) != 0)
        mps_fatal(c, __FILE__, __LINE__, "radix size %zu bytes is not in 1 .. %d", rsize, MPS_MAX_KEY_WORDS * 8);
    CUDA_OK(c, cudaSetDevice(c->device));
    if ((nmemb && is_device_pointer(c, base)) || (outnmemb && is_device_pointer(c, out)))
        mps_fatal(c, __FILE__, __LINE__, "a radix() callback runs on the host and cannot read device memory: "
                                          "pass host buffers, or describe the key with mpsort_mpi_newarray_desc");

    const size_t E2 = rpad + elsize;
    const int inplace = (out == base);
    /* in place (out == base) both views share one buffer: size it for the larger count */
    const size_t nin = (inplace && outnmemb > nmemb) ? outnmemb : nmemb;
    char * aug_in = (char *) mps_host_malloc("mpsort radix+record (in)", (nin ? nin : 1) * E2, __FILE__, __LINE__);
    char * aug_out = inplace ? aug_in
        : (char *) mps_host_malloc("mpsort radix+record (out)", (outnmemb ? outnmemb : 1) * E2, __FILE__, __LINE__);
    if (!aug_in || !aug_out) mps_fatal(c, __FILE__, __LINE__, "out of host memory for %zu + %zu augmented records of %zu bytes", nmemb, outnmemb, E2);
    mpsort_callback_pack(base, nmemb, elsize, radix, rsize, arg, aug_in);
    mpsort_mpi_newarray_desc_impl(aug_in, nmemb, aug_out, outnmemb, E2, &d, c, line, file);
    mpsort_callback_unpack(aug_out, outnmemb, elsize, rsize, out);
    if (!inplace) mps_host_free(aug_out, __FILE__, __LINE__);
    mps_host_free(aug_in, __FILE__, __LINE__);
}

void mpsort_mpi_impl(void * base, size_t nmemb, size_t elsize,
        mpsort_radix_func radix, size_t rsize, void * arg,
        mpsort_comm_t comm, const int line, const char * file)
{
    mpsort_mpi_newarray_impl(base, nmemb, base, nmemb, elsize, radix, rsize, arg, comm, line, file);
}

void radix_sort(void * base, size_t nmemb, size_t size,
        mpsort_radix_func radix, size_t rsize, void * arg)
{
    /* the reference's radix_sort has no communicator: the cached size-1 communicator of the
     * calling thread's current CUDA device */
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) { cudaGetLastError(); device = 0; }
    if (device < 0 || device >= 64) device = 0;
    mpsort_mpi_newarray_impl(base, nmemb, base, nmemb, size, radix, rsize, arg, cached_self_comm(device), __LINE__, __FILE__);
}
