/*
 * mpsort_comm.c -- the communicator of mpsort-b200: what replaces MPI_Comm and the
 * MPI calls of the reference (inventory: SURVEY.md 2.2, call sites C1-C14).
 *
 * Three transports behind one struct:
 *   SELF   one rank, no collective at all (single-GPU path)
 *   NCCL   one process per GPU; ncclAllReduce / ncclAllGather for counts, grouped
 *          ncclSend/ncclRecv over NVLink for the record exchange
 *          (replaces MPI_Alltoallv and MPI_Alltoallv_sparse, mp-mpiu.c:69-236)
 *   LOCAL  ranks are host threads of one process; records move with
 *          cudaMemcpyAsync (peer or same-device copies), counts through host memory
 */
#include <stdarg.h>
#include <string.h>
#include <unistd.h>

#include "mpsort_internal.h"

__thread const char * mps_caller_file = "?";
__thread int mps_caller_line = 0;

void mps_fatal(struct mpsort_comm * c, const char * file, int line, const char * fmt, ...)
{
    va_list va;
    fprintf(stderr, "MPSort: ");
    va_start(va, fmt);
    vfprintf(stderr, fmt, va);
    va_end(va);
    fprintf(stderr, " [%s:%d, rank %d]. Caller site: %s:%d\n",
            file, line, c ? c->rank : -1, mps_caller_file, mps_caller_line);
    fflush(stderr);
    if (c && c->kind == MPS_T_NCCL && c->nccl) {
        /* the moral equivalent of MPI_Abort(comm, -1): do not leave peers spinning */
        ncclCommAbort(c->nccl);
    }
    abort();
}

/* ------------------------------------------------------------------------- */
/* allocator hook (reference mp-mpiu.c:9-59) for host allocations             */

static void * default_malloc(const char * name, size_t size, const char * file, const int line, void * ud)
{ (void) name; (void) file; (void) line; (void) ud; return malloc(size); }
static void default_free(void * ptr, const char * file, const int line, void * ud)
{ (void) file; (void) line; (void) ud; free(ptr); }

static struct {
    mpiu_malloc_func malloc_func;
    mpiu_free_func free_func;
    void * userdata;
} g_mem = { default_malloc, default_free, NULL };

void mpiu_set_malloc(mpiu_malloc_func m, mpiu_free_func f, void * userdata)
{
    g_mem.malloc_func = m ? m : default_malloc;
    g_mem.free_func = f ? f : default_free;
    g_mem.userdata = userdata;
}

void * mps_host_malloc(const char * name, size_t size, const char * file, int line)
{
    return g_mem.malloc_func(name, size, file, line, g_mem.userdata);
}
void mps_host_free(void * ptr, const char * file, int line)
{
    g_mem.free_func(ptr, file, line, g_mem.userdata);
}

void MPIU_Set_verbose_malloc(mpsort_comm_t comm)
{
    if (comm) comm->verbose_malloc = 1;
}

/* ------------------------------------------------------------------------- */
/* arena                                                                      */

static const char * slot_names[MPS_NSLOTS] = {
    "din", "dout", "keywords", "keys_b", "keys_a", "idx_a", "idx_b", "sortedkeys",
    "hist", "lookback", "sendbuf", "recvbuf", "splitters", "stage", "stage2", "misc",
    "merge_samples", "merge_cuts", "merge_sorted_samples", "merge_sample_ids", "predictor"
};

void * mps_arena_get(struct mpsort_comm * c, int slot, size_t bytes)
{
    if (bytes == 0) bytes = 256;
    if (c->slot[slot].cap >= bytes) return c->slot[slot].ptr;
    if (c->slot[slot].ptr && slot == (c->p2p.pull ? MPS_S_SEND : MPS_S_RECV) && c->kind == MPS_T_NCCL && !c->p2p.disabled) {
        /* peers may still have this buffer mapped (CUDA IPC) and push into it by DMA: it is
         * NEVER freed before the communicator is destroyed (so its address cannot be handed
         * out again either). Growth is rare (grow-only, 25 % headroom). */
        CUDA_OK(c, cudaStreamSynchronize(c->stream));
        if (c->p2p.nzombies == c->p2p.zombies_cap) {
            const int cap = c->p2p.zombies_cap ? 2 * c->p2p.zombies_cap : 16;
            void ** z = (void **) realloc(c->p2p.zombies, sizeof(void *) * (size_t) cap);
            if (!z) mps_fatal(c, __FILE__, __LINE__, "out of host memory");
            c->p2p.zombies = z;
            c->p2p.zombies_cap = cap;
        }
        c->p2p.zombies[c->p2p.nzombies++] = c->slot[slot].ptr;
        c->p2p.generation++;
        c->slot[slot].ptr = NULL;
        c->slot[slot].cap = 0;
        bytes += bytes / 4;
    }
    if (c->slot[slot].ptr) {
        /* buffers may still be in use by work queued on the stream */
        CUDA_OK(c, cudaStreamSynchronize(c->stream));
        CUDA_OK(c, cudaFree(c->slot[slot].ptr));
        if (c->verbose_malloc)
            fprintf(stderr, "MPIU_Free: T%04d %16p : device arena '%s'\n", c->rank, c->slot[slot].ptr, slot_names[slot]);
        c->slot[slot].ptr = NULL;
        c->slot[slot].cap = 0;
    }
    /* round up to 2 MiB so that slowly growing inputs do not reallocate every call */
    size_t cap = (bytes + ((size_t) 2 << 20) - 1) & ~(((size_t) 2 << 20) - 1);
    void * p = NULL;
    cudaError_t e = cudaMalloc(&p, cap);
    if (e != cudaSuccess) {
        mps_fatal(c, __FILE__, __LINE__, "cannot allocate %zu bytes of device memory for '%s' (%s)",
                  cap, slot_names[slot], cudaGetErrorString(e));
    }
    if (c->verbose_malloc)
        fprintf(stderr, "MPIU_Malloc: T%04d %16p : device arena '%s' size = %zu\n", c->rank, p, slot_names[slot], cap);
    c->slot[slot].ptr = p;
    c->slot[slot].cap = cap;
    return p;
}

void mps_merge_ovf_begin(struct mpsort_comm * c)
{
    if (!c->d_merge_ovf) {
        CUDA_OK(c, cudaMalloc((void **) &c->d_merge_ovf, 256));
        CUDA_OK(c, cudaMallocHost((void **) &c->h_merge_ovf, 256));
    }
    c->h_merge_ovf[0] = c->h_merge_ovf[1] = 0;
    c->merge_ovf_pending = 0;
    CUDA_OK(c, cudaMemsetAsync(c->d_merge_ovf, 0, 2 * sizeof(uint32_t), c->stream));
}

void mps_merge_ovf_fetch(struct mpsort_comm * c)
{
    if (!c->d_merge_ovf) return;
    CUDA_OK(c, cudaMemcpyAsync(c->h_merge_ovf, c->d_merge_ovf, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    c->merge_ovf_pending = 1;
}

void mps_merge_ovf_check(struct mpsort_comm * c)
{
    if (!c->merge_ovf_pending) return;
    c->merge_ovf_pending = 0;
    if (c->h_merge_ovf[0] != 0)
        mps_fatal(c, __FILE__, __LINE__, "serious bug: %u merge tiles exceeded their bound", c->h_merge_ovf[0]);
}

void * mps_host_stage(struct mpsort_comm * c, size_t bytes)
{
    if (c->h_stage_cap >= bytes) return c->h_stage;
    if (c->h_stage) {
        CUDA_OK(c, cudaStreamSynchronize(c->stream));   /* async copies may still read it */
        CUDA_OK(c, cudaFreeHost(c->h_stage));
    }
    size_t cap = bytes < 65536 ? 65536 : bytes;
    CUDA_OK(c, cudaMallocHost(&c->h_stage, cap));
    c->h_stage_cap = cap;
    return c->h_stage;
}

/* ------------------------------------------------------------------------- */
/* per-kernel-class timing                                                    */

void mps_kt_begin(struct mpsort_comm * c, int cls)
{
    struct mps_ktimes * k = &c->kt;
    if (!k->on) return;
    if (k->n >= MPS_KT_MAX) { CUDA_OK(c, cudaStreamSynchronize(c->stream)); mps_kt_collect(c); }
    while (k->nev < 2 * (k->n + 1)) { CUDA_OK(c, cudaEventCreate(&k->ev[k->nev])); k->nev++; }
    k->cls[k->n] = k->force_cls > 0 ? k->force_cls : cls;
    CUDA_OK(c, cudaEventRecord(k->ev[2 * k->n], c->stream));
}

void mps_kt_end(struct mpsort_comm * c)
{
    struct mps_ktimes * k = &c->kt;
    if (!k->on) return;
    CUDA_OK(c, cudaEventRecord(k->ev[2 * k->n + 1], c->stream));
    k->n++;
}

void mps_kt_collect(struct mpsort_comm * c)
{
    struct mps_ktimes * k = &c->kt;
    int i;
    for (i = 0; i < k->n; i++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, k->ev[2 * i], k->ev[2 * i + 1]) == cudaSuccess) {
            k->ms[k->cls[i]] += ms;
            k->launches[k->cls[i]]++;
        } else cudaGetLastError();
    }
    k->n = 0;
}

/* ------------------------------------------------------------------------- */
/* construction                                                               */

static struct mpsort_comm * comm_alloc(int kind, int rank, int size, int device)
{
    struct mpsort_comm * c = (struct mpsort_comm *) calloc(1, sizeof(*c));
    if (!c) { fprintf(stderr, "MPSort: out of host memory\n"); abort(); }
    c->kind = kind; c->rank = rank; c->size = size; c->device = device;
    if (size > MPS_MAX_RANKS) mps_fatal(c, __FILE__, __LINE__, "communicator size %d exceeds the supported maximum %d", size, MPS_MAX_RANKS);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        mps_fatal(c, __FILE__, __LINE__,
                  "no CUDA device available (%s): mpsort-b200 has no CPU fallback",
                  cudaGetErrorString(e));
    }
    if (device < 0 || device >= ndev) mps_fatal(c, __FILE__, __LINE__, "device %d out of range (have %d)", device, ndev);
    CUDA_OK(c, cudaSetDevice(device));
    {
        /* the payload gather reads one random record per thread: fetch only the
         * sectors asked for instead of whole 128-byte lines */
        const char * g = getenv("MPSORT_L2_FETCH_GRANULARITY");
        const int gran = g ? atoi(g) : 0;
        if (gran > 0) {
            cudaError_t e2 = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t) gran);
            if (e2 != cudaSuccess) cudaGetLastError();
        }
    }
    CUDA_OK(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    return c;
}

int mpsort_comm_get_unique_id(void * id)
{
    ncclUniqueId uid;
    if (sizeof(uid) > MPSORT_UNIQUE_ID_BYTES) return -1;
    ncclResult_t r = ncclGetUniqueId(&uid);
    if (r != ncclSuccess) {
        fprintf(stderr, "MPSort: ncclGetUniqueId failed: %s\n", ncclGetErrorString(r));
        return (int) r;
    }
    memset(id, 0, MPSORT_UNIQUE_ID_BYTES);
    memcpy(id, &uid, sizeof(uid));
    return 0;
}

mpsort_comm_t mpsort_comm_init_rank(int rank, int size, const void * unique_id, int device)
{
    if (size <= 1) return mpsort_comm_self(device);
    struct mpsort_comm * c = comm_alloc(MPS_T_NCCL, rank, size, device);
    ncclUniqueId uid;
    memcpy(&uid, unique_id, sizeof(uid));
    NCCL_OK(c, ncclCommInitRank(&c->nccl, size, uid, rank));
    /* Exchange transport inside one box (profiles/r01_p2p_exchange.log, GB/s per GPU and
     * direction at 2 / 8 GPUs): peers' receive buffers are mapped with CUDA IPC and every
     * slice moves by ONE cudaMemcpyAsync at a time in shifted order me+1, me+2, ... (at every
     * step the pairs form a permutation): 768 / 749, and no SM is busy, so the merge of an
     * earlier part can overlap. NCCL grouped send/recv: 543 / 649; peer stores from a copy
     * kernel: 634 / 557; seven DMA copies at once: - / 442.
     * MPSORT_NO_P2P=1: NCCL send/recv. MPSORT_P2P_CE=0: the copy kernel; =k: k DMA copies in flight. */
    c->p2p.disabled = getenv("MPSORT_NO_P2P") ? 1 : 0;
    c->p2p.pull = getenv("MPSORT_P2P_PULL") ? 1 : 0;
    c->p2p.copy_engine = getenv("MPSORT_P2P_CE") ? atoi(getenv("MPSORT_P2P_CE")) : 1;
    if (c->p2p.copy_engine < 0) c->p2p.copy_engine = 0;
    if (c->p2p.copy_engine > 7) c->p2p.copy_engine = 7;
    return c;
}

mpsort_comm_t mpsort_comm_self(int device)
{
    return comm_alloc(MPS_T_SELF, 0, 1, device);
}

int mpsort_comm_init_local_group(int size, const int * devices, mpsort_comm_t * comms)
{
    int i, j;
    if (size < 1 || size > MPS_MAX_RANKS) return -1;
    if (size == 1) { comms[0] = mpsort_comm_self(devices[0]); return 0; }
    struct mps_local_group * g = (struct mps_local_group *) calloc(1, sizeof(*g));
    if (!g) return -2;
    g->size = size;
    g->refcount = size;
    pthread_barrier_init(&g->barrier, NULL, (unsigned) size);
    pthread_mutex_init(&g->lock, NULL);
    for (i = 0; i < size; i++) {
        comms[i] = comm_alloc(MPS_T_LOCAL, i, size, devices[i]);
        comms[i]->grp = g;
    }
    /* kernels of one rank read buffers of the others (the in-process all-reduce, the pull
     * exchange): peer access between distinct devices is a REQUIREMENT of this transport, so a pair
     * that cannot have it is an error here, at construction, not an illegal-address fault in the
     * middle of a sort (same-device ranks need nothing) */
    for (i = 0; i < size; i++) {
        for (j = 0; j < size; j++) {
            if (devices[i] == devices[j]) continue;
            int can = 0;
            cudaError_t e = cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
            if (e == cudaSuccess && can) {
                cudaSetDevice(devices[i]);
                e = cudaDeviceEnablePeerAccess(devices[j], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            } else if (e == cudaSuccess) e = cudaErrorPeerAccessUnsupported;
            if (e != cudaSuccess) {
                fprintf(stderr, "MPSort: devices %d and %d of an in-process group cannot access each other's memory (%s); "
                                "use one process per GPU (mpsort_comm_init_rank)\n", devices[i], devices[j], cudaGetErrorString(e));
                cudaGetLastError();
                for (i = 0; i < size; i++) { comms[i]->grp = NULL; comms[i]->kind = MPS_T_SELF; mpsort_comm_destroy(comms[i]); comms[i] = NULL; }
                pthread_barrier_destroy(&g->barrier);
                pthread_mutex_destroy(&g->lock);
                free(g);
                return -3;
            }
        }
    }
    return 0;
}

void mpsort_comm_destroy(mpsort_comm_t c)
{
    int s;
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (s = 0; s < c->size && c->kind == MPS_T_NCCL; s++)
        if (c->p2p.peer_base[s]) cudaIpcCloseMemHandle(c->p2p.peer_base[s]);
    if (c->kind == MPS_T_NCCL && c->peer.state > 0)
        for (s = 0; s < c->size; s++) if (s != c->rank && c->peer.box[s]) cudaIpcCloseMemHandle(c->peer.box[s]);
    if (c->kind == MPS_T_NCCL && c->size > 1) mpsort_comm_barrier(c);   /* everyone unmapped before anyone frees */
    if (c->peer.mine) cudaFree(c->peer.mine);
    for (s = 0; s < c->p2p.nzombies; s++) cudaFree(c->p2p.zombies[s]);
    free(c->p2p.zombies);
    for (s = 0; s < MPS_NSLOTS; s++) if (c->slot[s].ptr) cudaFree(c->slot[s].ptr);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->d_merge_ovf) cudaFree(c->d_merge_ovf);
    if (c->h_merge_ovf) cudaFreeHost(c->h_merge_ovf);
    if (c->p2p.d_flag) cudaFree(c->p2p.d_flag);
    if (c->p2p.ce_created) {
        for (s = 0; s < 8; s++) cudaStreamDestroy(c->p2p.ce_stream[s]);
        for (s = 0; s < 10; s++) cudaEventDestroy(c->p2p.ce_ev[s]);
    }
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->io_created) {
        for (s = 0; s < MPS_MAX_CHUNKS; s++) cudaEventDestroy(c->in.ev[s]);
        cudaEventDestroy(c->io_ev[0]); cudaEventDestroy(c->io_ev[1]);
        cudaStreamDestroy(c->h2d_stream); cudaStreamDestroy(c->d2h_stream);
    }
    if (c->phase_ev_created) { int i; for (i = 0; i <= MPS_MAX_RANKS; i++) cudaEventDestroy(c->phase_ev[i]); }
    if (c->kind == MPS_T_NCCL && c->nccl) ncclCommDestroy(c->nccl);
    if (c->kind == MPS_T_LOCAL && c->grp) {
        int last;
        pthread_mutex_lock(&c->grp->lock);
        last = (--c->grp->refcount == 0);
        pthread_mutex_unlock(&c->grp->lock);
        if (last) {
            pthread_barrier_destroy(&c->grp->barrier);
            pthread_mutex_destroy(&c->grp->lock);
            free(c->grp);
        }
    }
    { int i; for (i = 0; i < c->kt.nev; i++) cudaEventDestroy(c->kt.ev[i]); }
    if (c->timers.created) { int i; for (i = 0; i < MPS_MAX_TIMERS; i++) cudaEventDestroy(c->timers.ev[i]); }
    cudaStreamDestroy(c->stream);
    free(c);
}

int mpsort_comm_rank(mpsort_comm_t c) { return c->rank; }
int mpsort_comm_size(mpsort_comm_t c) { return c->size; }
int mpsort_comm_device(mpsort_comm_t c) { return c->device; }
void * mpsort_comm_stream(mpsort_comm_t c) { return (void *) c->stream; }

/* ------------------------------------------------------------------------- */
/* small host collectives                                                     */

static void local_barrier(struct mpsort_comm * c)
{
    pthread_barrier_wait(&c->grp->barrier);
}

void mpsort_comm_allgatherv_host(mpsort_comm_t c, const void * send, size_t nbytes,
                                 void * recv, const size_t * recvcounts)
{
    int j;
    size_t total = 0, myoff = 0;
    for (j = 0; j < c->size; j++) {
        if (j == c->rank) myoff = total;
        total += recvcounts[j];
    }
    if (recvcounts[c->rank] != nbytes) mps_fatal(c, __FILE__, __LINE__, "allgatherv: inconsistent counts");
    if (c->kind == MPS_T_SELF) {
        if (nbytes) memmove(recv, send, nbytes);
        return;
    }
    if (c->kind == MPS_T_LOCAL) {
        c->grp->slot[c->rank] = send;
        local_barrier(c);
        size_t off = 0;
        for (j = 0; j < c->size; j++) {
            if (recvcounts[j]) memcpy((char *) recv + off, c->grp->slot[j], recvcounts[j]);
            off += recvcounts[j];
        }
        local_barrier(c);
        return;
    }
    /* NCCL: stage through device memory; every rank broadcasts its piece (grouped) */
    CUDA_OK(c, cudaSetDevice(c->device));
    if (total == 0) return;
    char * h = (char *) mps_host_stage(c, total);
    char * d = (char *) mps_arena_get(c, MPS_S_STAGE, total);
    if (nbytes) {
        memcpy(h + myoff, send, nbytes);
        CUDA_OK(c, cudaMemcpyAsync(d + myoff, h + myoff, nbytes, cudaMemcpyHostToDevice, c->stream));
    }
    NCCL_OK(c, ncclGroupStart());
    size_t off = 0;
    for (j = 0; j < c->size; j++) {
        if (recvcounts[j])
            NCCL_OK(c, ncclBroadcast(d + off, d + off, recvcounts[j], ncclUint8, j, c->nccl, c->stream));
        off += recvcounts[j];
    }
    NCCL_OK(c, ncclGroupEnd());
    CUDA_OK(c, cudaMemcpyAsync(h, d, total, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    memcpy(recv, h, total);
}

void mpsort_comm_allgather_host(mpsort_comm_t c, const void * send, void * recv, size_t nbytes)
{
    int j;
    if (c->kind == MPS_T_NCCL) {
        /* equal pieces: one ncclAllGather (replaces the MPI_Allgathers of
         * mpsort-mpi.c:633-640 and mp-mpiu.c:415-416) */
        CUDA_OK(c, cudaSetDevice(c->device));
        const size_t total = nbytes * (size_t) c->size;
        if (total == 0) return;
        char * h = (char *) mps_host_stage(c, total);
        char * d = (char *) mps_arena_get(c, MPS_S_STAGE, total);
        memcpy(h + nbytes * c->rank, send, nbytes);
        CUDA_OK(c, cudaMemcpyAsync(d + nbytes * c->rank, h + nbytes * c->rank, nbytes, cudaMemcpyHostToDevice, c->stream));
        NCCL_OK(c, ncclAllGather(d + nbytes * c->rank, d, nbytes, ncclUint8, c->nccl, c->stream));
        CUDA_OK(c, cudaMemcpyAsync(h, d, total, cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(c, cudaStreamSynchronize(c->stream));
        memcpy(recv, h, total);
        return;
    }
    size_t counts[MPS_MAX_RANKS];
    for (j = 0; j < c->size; j++) counts[j] = nbytes;
    mpsort_comm_allgatherv_host(c, send, nbytes, recv, counts);
}

void mpsort_comm_barrier(mpsort_comm_t c)
{
    if (c->kind == MPS_T_SELF) return;
    if (c->kind == MPS_T_LOCAL) { local_barrier(c); return; }
    char token = 1, all[MPS_MAX_RANKS];
    mpsort_comm_allgather_host(c, &token, all, 1);
}

/* ------------------------------------------------------------------------- */
/* device collectives                                                         */

/* in-place sum of count u64 on the device, stream ordered.
 * Replaces the two MPI_Allreduce per bisection round (mpsort-mpi.c:391-394). */
void mps_comm_allreduce_u64_dev(struct mpsort_comm * c, uint64_t * dptr, size_t count)
{
    if (c->kind == MPS_T_SELF || count == 0) return;
    if (c->kind == MPS_T_NCCL) {
        NCCL_OK(c, ncclAllReduce(dptr, dptr, count, ncclUint64, ncclSum, c->nccl, c->stream));
        return;
    }
    /* LOCAL: every rank's kernel sums all ranks' buffers into a private scratch */
    const uint64_t * srcs[MPS_MAX_RANKS];
    int j;
    uint64_t * tmp = (uint64_t *) mps_arena_get(c, MPS_S_STAGE2, count * sizeof(uint64_t));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    c->grp->slot[c->rank] = dptr;
    local_barrier(c);
    for (j = 0; j < c->size; j++) srcs[j] = (const uint64_t *) c->grp->slot[j];
    KERN_OK(c, mpsk_sum_u64(tmp, srcs, c->size, count, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    local_barrier(c);   /* everyone has read every dptr: now it may be overwritten */
    CUDA_OK(c, cudaMemcpyAsync(dptr, tmp, count * sizeof(uint64_t), cudaMemcpyDeviceToDevice, c->stream));
}

/*
 * The record exchange (one phase of it). The receive buffer is laid out in source-rank
 * order exactly like the reference's RecvDispl (mpsort-mpi.c:490-503).
 *
 * dense == 0 (AUTO / REQUIRE_SPARSE): zero-length pairs are skipped, the behaviour
 *   of MPI_Alltoallv_sparse (mp-mpiu.c:184-214). Safe with NCCL because both sides
 *   of every pair hold the same count matrix.
 * dense != 0 (DISABLE_SPARSE): every pair is posted, like MPI_Alltoallv (mp-mpiu.c:140).
 */
static void exchange_p2p(struct mpsort_comm * c, const void * sendbuf, const int64_t * sendoff, const int64_t * sendcnt,
        void * recvbuf, int64_t recvoff, const int64_t * recvcnt,
        const int64_t * peer_recvoff, const int64_t * peer_sendoff, size_t elsize, uint64_t * bytes_remote);

void mps_comm_exchange(struct mpsort_comm * c, const void * sendbuf, const int64_t * sendoff, const int64_t * sendcnt,
        void * recvbuf, int64_t recvoff, const int64_t * recvcnt,
        const int64_t * peer_recvoff, const int64_t * peer_sendoff,
        size_t elsize, int dense, int use_p2p, uint64_t * bytes_remote)
{
    const int p = c->size, me = c->rank;
    int j, k;
    uint64_t remote = 0;
    int64_t rdispl[MPS_MAX_RANKS + 1];
    rdispl[0] = recvoff;
    for (j = 0; j < p; j++) rdispl[j + 1] = rdispl[j] + recvcnt[j];

    if (c->kind == MPS_T_SELF) {
        if (sendcnt[0] > 0 && (const char *) sendbuf + (size_t) sendoff[0] * elsize != (char *) recvbuf + (size_t) recvoff * elsize)
            CUDA_OK(c, cudaMemcpyAsync((char *) recvbuf + (size_t) recvoff * elsize, (const char *) sendbuf + (size_t) sendoff[0] * elsize,
                                       (size_t) sendcnt[0] * elsize, cudaMemcpyDeviceToDevice, c->stream));
        return;
    }
    if (c->kind == MPS_T_NCCL && use_p2p) {
        exchange_p2p(c, sendbuf, sendoff, sendcnt, recvbuf, recvoff, recvcnt, peer_recvoff, peer_sendoff, elsize, bytes_remote);
        return;
    }
    if (c->kind == MPS_T_NCCL) {
        NCCL_OK(c, ncclGroupStart());
        for (k = 0; k < p; k++) {
            if (k == me) continue;
            if (sendcnt[k] > 0 || dense) {
                NCCL_OK(c, ncclSend((const char *) sendbuf + (size_t) sendoff[k] * elsize,
                                    (size_t) sendcnt[k] * elsize, ncclUint8, k, c->nccl, c->stream));
                remote += (uint64_t) sendcnt[k] * elsize;
            }
            if (recvcnt[k] > 0 || dense) {
                NCCL_OK(c, ncclRecv((char *) recvbuf + (size_t) rdispl[k] * elsize,
                                    (size_t) recvcnt[k] * elsize, ncclUint8, k, c->nccl, c->stream));
            }
        }
        NCCL_OK(c, ncclGroupEnd());
        if (sendcnt[me] > 0 && !c->p2p.skip_self)
            CUDA_OK(c, cudaMemcpyAsync((char *) recvbuf + (size_t) rdispl[me] * elsize,
                                       (const char *) sendbuf + (size_t) sendoff[me] * elsize,
                                       (size_t) sendcnt[me] * elsize, cudaMemcpyDeviceToDevice, c->stream));
        if (bytes_remote) *bytes_remote += remote;
        return;
    }
    /* LOCAL: pull from every source's send buffer */
    CUDA_OK(c, cudaStreamSynchronize(c->stream));   /* my sendbuf is complete */
    c->grp->slot2[me] = sendbuf;
    local_barrier(c);
    for (j = 0; j < p; j++) {
        if (recvcnt[j] <= 0 || (j == me && c->p2p.skip_self)) continue;
        const char * src = (const char *) c->grp->slot2[j] + (size_t) peer_sendoff[j] * elsize;
        CUDA_OK(c, cudaMemcpyAsync((char *) recvbuf + (size_t) rdispl[j] * elsize, src,
                                   (size_t) recvcnt[j] * elsize, cudaMemcpyDefault, c->stream));
    }
    for (k = 0; k < p; k++) if (k != me) remote += (uint64_t) sendcnt[k] * elsize;
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    local_barrier(c);   /* sources may reuse their send buffers */
    if (bytes_remote) *bytes_remote += remote;
}

/* whole exchange from the full cut matrix: rank j's sorted records [cut[j][k], cut[j][k+1])
 * go to rank k (used by the small-input gather path) */
void mps_comm_alltoallv_dev(struct mpsort_comm * c, const void * sendbuf, void * recvbuf,
        const int64_t * cut, size_t elsize, int dense, uint64_t * bytes_remote)
{
    const int p = c->size, me = c->rank;
    int j, k;
    int64_t sendoff[MPS_MAX_RANKS], sendcnt[MPS_MAX_RANKS], recvcnt[MPS_MAX_RANKS];
    int64_t peer_recvoff[MPS_MAX_RANKS], peer_sendoff[MPS_MAX_RANKS];
#define CUT(j, k) cut[(size_t) (j) * (p + 1) + (k)]
    for (k = 0; k < p; k++) {
        sendoff[k] = CUT(me, k);
        sendcnt[k] = CUT(me, k + 1) - CUT(me, k);
        recvcnt[k] = CUT(k, me + 1) - CUT(k, me);
        peer_sendoff[k] = CUT(k, me);
        peer_recvoff[k] = 0;
        for (j = 0; j < me; j++) peer_recvoff[k] += CUT(j, k + 1) - CUT(j, k);
    }
#undef CUT
    if (bytes_remote) *bytes_remote = 0;
    mps_comm_exchange(c, sendbuf, sendoff, sendcnt, recvbuf, 0, recvcnt, peer_recvoff, peer_sendoff, elsize, dense, 0, bytes_remote);
}

/* ------------------------------------------------------------------------- */
/* peer-store exchange                                                        */

void mps_comm_recv_info(struct mpsort_comm * c, void * recvbuf, void * sendbuf, struct mps_recv_info * info)
{
    memset(info, 0, sizeof(*info));
    if (c->kind != MPS_T_NCCL || c->p2p.disabled) return;
    /* pull: peers read my send buffer; push: peers write my receive buffer */
    void * buf = c->p2p.pull ? sendbuf : recvbuf;
    info->ptr = (uint64_t) (uintptr_t) buf;
    info->cap = c->slot[c->p2p.pull ? MPS_S_SEND : MPS_S_RECV].cap;
    info->generation = c->p2p.generation;
    cudaIpcMemHandle_t hdl;
    if (cudaIpcGetMemHandle(&hdl, buf) != cudaSuccess) {
        cudaGetLastError();
        info->ptr = 0;               /* tells everyone: no peer stores */
        return;
    }
    memcpy(info->handle, &hdl, sizeof(hdl) <= sizeof(info->handle) ? sizeof(hdl) : sizeof(info->handle));
}

int mps_comm_p2p_prepare(struct mpsort_comm * c, const struct mps_recv_info * all)
{
    int j, changed = 0, ok = 1;
    if (c->kind != MPS_T_NCCL || c->p2p.disabled) return 0;
    for (j = 0; j < c->size; j++) if (all[j].ptr == 0) ok = 0;       /* same table on every rank */
    if (!ok) { c->p2p.disabled = 1; return 0; }
    for (j = 0; j < c->size; j++) {
        /* my own entry counts too: every rank must reach the same verdict */
        if (c->p2p.peer_ptr[j] == all[j].ptr && c->p2p.peer_gen[j] == all[j].generation) continue;
        changed = 1;
    }
    /* every rank sees the same table and has the same history: `changed` agrees everywhere */
    if (!changed) return 1;
    int mine_ok = 1;
    c->p2p.peer_ptr[c->rank] = all[c->rank].ptr;
    c->p2p.peer_gen[c->rank] = all[c->rank].generation;
    for (j = 0; j < c->size; j++) {
        if (j == c->rank) continue;
        if (c->p2p.peer_ptr[j] == all[j].ptr && c->p2p.peer_gen[j] == all[j].generation && c->p2p.peer_base[j]) continue;
        if (c->p2p.peer_base[j]) { cudaIpcCloseMemHandle(c->p2p.peer_base[j]); c->p2p.peer_base[j] = NULL; }
        cudaIpcMemHandle_t hdl;
        memcpy(&hdl, all[j].handle, sizeof(hdl));
        void * p = NULL;
        if (cudaIpcOpenMemHandle(&p, hdl, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            mine_ok = 0;
            break;
        }
        c->p2p.peer_base[j] = p;
        c->p2p.peer_ptr[j] = all[j].ptr;
        c->p2p.peer_gen[j] = all[j].generation;
    }
    /* collective agreement: one rank that cannot map switches everyone back to NCCL */
    char flag = (char) mine_ok, flags[MPS_MAX_RANKS];
    mpsort_comm_allgather_host(c, &flag, flags, 1);
    for (j = 0; j < c->size; j++) if (!flags[j]) ok = 0;
    if (!ok) {
        for (j = 0; j < c->size; j++) if (c->p2p.peer_base[j]) { cudaIpcCloseMemHandle(c->p2p.peer_base[j]); c->p2p.peer_base[j] = NULL; }
        c->p2p.disabled = 1;
        if (c->rank == 0) fprintf(stderr, "MPSort: CUDA IPC mapping of peer buffers failed; exchanging with ncclSend/ncclRecv\n");
        return 0;
    }
    return 1;
}

static void exchange_p2p(struct mpsort_comm * c, const void * sendbuf, const int64_t * sendoff, const int64_t * sendcnt,
        void * recvbuf, int64_t recvoff, const int64_t * recvcnt,
        const int64_t * peer_recvoff, const int64_t * peer_sendoff, size_t elsize, uint64_t * bytes_remote)
{
    const int p = c->size, me = c->rank;
    int k;
    const void * src[MPS_MAX_RANKS];
    void * dst[MPS_MAX_RANKS];
    uint64_t bytes[MPS_MAX_RANKS], remote = 0;
    unsigned char isremote[MPS_MAX_RANKS];
    if (!c->p2p.d_flag) {
        CUDA_OK(c, cudaMalloc((void **) &c->p2p.d_flag, 256));
        CUDA_OK(c, cudaMemsetAsync(c->p2p.d_flag, 0, 256, c->stream));
    }
    if (c->p2p.pull) {
        /* every rank's send buffer must be complete before anyone reads it */
        NCCL_OK(c, ncclAllReduce(c->p2p.d_flag, c->p2p.d_flag, 1, ncclInt32, ncclSum, c->nccl, c->stream));
        int64_t rd = recvoff;
        for (k = 0; k < p; k++) {
            src[k] = (const char *) (k == me ? sendbuf : c->p2p.peer_base[k]) + (size_t) peer_sendoff[k] * elsize;
            dst[k] = (char *) recvbuf + (size_t) rd * elsize;
            bytes[k] = (uint64_t) recvcnt[k] * elsize;
            isremote[k] = (unsigned char) (k != me);
            rd += recvcnt[k];
            if (k != me) remote += (uint64_t) sendcnt[k] * elsize;
        }
    } else {
        for (k = 0; k < p; k++) {
            src[k] = (const char *) sendbuf + (size_t) sendoff[k] * elsize;
            dst[k] = (char *) (k == me ? recvbuf : c->p2p.peer_base[k]) + (size_t) peer_recvoff[k] * elsize;
            bytes[k] = (uint64_t) sendcnt[k] * elsize;
            isremote[k] = (unsigned char) (k != me);
            if (k != me) remote += bytes[k];
        }
    }
    if (c->p2p.skip_self) bytes[me] = 0;     /* the merge reads my own slice where it is */
    {
        /* segments in shifted order: me+1, me+2, ..., me (self last) */
        const void * rsrc[MPS_MAX_RANKS];
        void * rdst[MPS_MAX_RANKS];
        uint64_t rbytes[MPS_MAX_RANKS];
        unsigned char rrem[MPS_MAX_RANKS];
        for (k = 0; k < p; k++) {
            const int q = (me + 1 + k) % p;
            rsrc[k] = src[q]; rdst[k] = dst[q]; rbytes[k] = bytes[q]; rrem[k] = isremote[q];
        }
        if (c->p2p.copy_engine) {
            /* the slices move by DMA: no SM is busy with the exchange, so a merge of an
             * earlier part (MPSORT_EXCHANGE_PHASES) can have the whole GPU meanwhile */
            if (!c->p2p.ce_created) {
                for (k = 0; k < 8; k++) CUDA_OK(c, cudaStreamCreateWithFlags(&c->p2p.ce_stream[k], cudaStreamNonBlocking));
                for (k = 0; k < 10; k++) CUDA_OK(c, cudaEventCreateWithFlags(&c->p2p.ce_ev[k], cudaEventDisableTiming));
                c->p2p.ce_created = 1;
            }
            /* remote slices in shifted order (me+1, me+2, ...: at every step the pairs form a
             * permutation) dealt round-robin to copy_engine streams; my own slice on stream 7.
             * Every copy stream that gets a copy first waits for the send buffer (the main stream
             * up to here) and the main stream then waits for every one of them -- stream 7 included,
             * whatever p is (it used to be left out for p < 8: found by the CPU stream model,
             * tests/native/mock_async.cpp). (Cutting a slice into two or four copies that move at the same time
             * was measured at two GPUs, where one copy per direction is all there is: no gain,
             * profiles/r02_call_n2_nccl_parity_parts_candidates.log -- a single copy already runs at 770 GB/s.) */
            /* (a sparse exchange moves small slices: one at a time would be all launch latency, so they go out
             * on seven streams at once) */
            const int lanes = c->p2p.burst ? 7 : c->p2p.copy_engine;
            unsigned used = 0;
            for (k = 0; k < p; k++) if (rbytes[k]) used |= 1u << (rrem[k] ? k % lanes : 7);
            /* What the copies of this part wait for (the gate). The send buffer was complete, and every rank
             * past its local sort, before the FIRST part of an exchange, and the parts land in disjoint slices
             * of the receive buffers, so nothing but the first part's gate is NEEDED; what the gate decides is
             * how far the ranks may drift apart:
             *   chained 0  the main stream as it is now, i.e. behind the completion barrier of the part before:
             *              every part starts in lock step, and the shifted order keeps the pairs a permutation;
             *              the links idle for that barrier (~0.1 ms) after every part;
             *   chained 1  the gate of the first part: the copies of all parts queue up back to back. Two GPUs
             *              (one peer, nothing to collide with): 13.49 -> 13.13 ms. Eight GPUs: the ranks drift,
             *              several senders meet at one receiver and the exchange takes 6.9 ms instead of 5.95
             *              (profiles/r02_call_n2c_*.log, r02_call_n8_c_*.log);
             *   chained 2  the main stream as it was when the part BEFORE started (behind the barrier of the part
             *              before that): the links stay busy during a barrier, the drift is bounded by one part. */
            cudaEvent_t gate = c->p2p.ce_ev[8];
            if (c->p2p.chained != 1) {
                cudaEvent_t now = c->p2p.ce_ev[8 + (c->p2p.part & 1)];
                CUDA_OK(c, cudaEventRecord(now, c->stream));
                gate = (c->p2p.chained == 2 && c->p2p.part > 0) ? c->p2p.ce_ev[8 + ((c->p2p.part - 1) & 1)] : now;
            }
            for (k = 0; k < 8; k++) if ((used >> k) & 1u) CUDA_OK(c, cudaStreamWaitEvent(c->p2p.ce_stream[k], gate, 0));
            for (k = 0; k < p; k++)
                if (rbytes[k])
                    CUDA_OK(c, cudaMemcpyAsync(rdst[k], rsrc[k], (size_t) rbytes[k], cudaMemcpyDeviceToDevice,
                                               c->p2p.ce_stream[rrem[k] ? k % lanes : 7]));
            for (k = 0; k < 8; k++) {
                if (!((used >> k) & 1u)) continue;
                CUDA_OK(c, cudaEventRecord(c->p2p.ce_ev[k], c->p2p.ce_stream[k]));
                CUDA_OK(c, cudaStreamWaitEvent(c->stream, c->p2p.ce_ev[k], 0));
            }
        } else {
            KERN_OK(c, mpsk_p2p_alltoallv(rsrc, rdst, rbytes, rrem, p, c->stream));
        }
    }
    /* push: all stores into my buffer are complete when everyone's kernel is; pull: nobody
     * may reuse its send buffer before everyone has read it. A one-word all-reduce on the
     * stream is the barrier (stream ordered, no host involvement). A caller that moves several
     * parts in one step (a sparse exchange) asks for it after the last part only. */
    if (!c->p2p.skip_barrier)
        NCCL_OK(c, ncclAllReduce(c->p2p.d_flag, c->p2p.d_flag, 1, ncclInt32, ncclSum, c->nccl, c->stream));
    if (bytes_remote) *bytes_remote += remote;
}

/* CANDIDATE (MPSORT_FUSED_PACK=1, push transport only): one phase of the exchange straight from the
 * unsorted records `base` and the sorted permutation `idx`: my records at sorted positions
 * [sendoff[k], sendoff[k] + sendcnt[k]) go to item peer_recvoff[k] of rank k's receive buffer.
 * Same completion barrier as exchange_p2p. Needs a successful mps_comm_p2p_prepare. */
void mps_comm_exchange_gather(struct mpsort_comm * c, const void * base, const uint32_t * idx,
        const int64_t * sendoff, const int64_t * sendcnt, void * recvbuf, const int64_t * peer_recvoff,
        size_t elsize, uint64_t * bytes_remote)
{
    const int p = c->size, me = c->rank;
    int k;
    const uint32_t * sidx[MPS_MAX_RANKS];
    void * dst[MPS_MAX_RANKS];
    uint64_t nrec[MPS_MAX_RANKS], remote = 0;
    if (!c->p2p.d_flag) {
        CUDA_OK(c, cudaMalloc((void **) &c->p2p.d_flag, 256));
        CUDA_OK(c, cudaMemsetAsync(c->p2p.d_flag, 0, 256, c->stream));
    }
    /* segments in shifted order: me+1, me+2, ..., me (self last) */
    for (k = 0; k < p; k++) {
        const int q = (me + 1 + k) % p;
        sidx[k] = idx + sendoff[q];
        dst[k] = (char *) (q == me ? recvbuf : c->p2p.peer_base[q]) + (size_t) peer_recvoff[q] * elsize;
        nrec[k] = (uint64_t) sendcnt[q];
        if (q != me) remote += (uint64_t) sendcnt[q] * elsize;
    }
    KERN_OK(c, mpsk_p2p_gather_alltoallv(base, sidx, dst, nrec, elsize, p, c->stream));
    if (!c->p2p.skip_barrier)
        NCCL_OK(c, ncclAllReduce(c->p2p.d_flag, c->p2p.d_flag, 1, ncclInt32, ncclSum, c->nccl, c->stream));
    if (bytes_remote) *bytes_remote += remote;
}

/* ------------------------------------------------------------------------- */
/* CANDIDATE (MPSORT_PEER_SPLITTER=1): mailboxes of the one-kernel splitter descent */

int mps_comm_peer_boxes_prepare(struct mpsort_comm * c)
{
    int j, ok = 1;
    if (c->peer.state) return c->peer.state > 0;
    if (c->kind == MPS_T_SELF || c->size > MPS_MAX_RANKS) { c->peer.state = -1; return 0; }
    const size_t bytes = mpsk_peer_box_bytes();
    CUDA_OK(c, cudaSetDevice(c->device));
    CUDA_OK(c, cudaMalloc(&c->peer.mine, bytes + 256));
    CUDA_OK(c, cudaMemsetAsync(c->peer.mine, 0, bytes + 256, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    c->peer.box[c->rank] = c->peer.mine;
    if (c->kind == MPS_T_LOCAL) {
        c->grp->slot[c->rank] = c->peer.mine;
        local_barrier(c);
        for (j = 0; j < c->size; j++) c->peer.box[j] = (void *) c->grp->slot[j];
        local_barrier(c);
        c->peer.state = 1;
        return 1;
    }
    struct { int ok; unsigned char handle[64]; } mine, all[MPS_MAX_RANKS];
    memset(&mine, 0, sizeof(mine));
    cudaIpcMemHandle_t hdl;
    if (cudaIpcGetMemHandle(&hdl, c->peer.mine) == cudaSuccess) {
        mine.ok = 1;
        memcpy(mine.handle, &hdl, sizeof(hdl) <= sizeof(mine.handle) ? sizeof(hdl) : sizeof(mine.handle));
    } else cudaGetLastError();
    mpsort_comm_allgather_host(c, &mine, all, sizeof(mine));
    for (j = 0; j < c->size; j++) if (!all[j].ok) ok = 0;
    int mine_ok = ok;
    for (j = 0; j < c->size && mine_ok; j++) {
        if (j == c->rank) continue;
        void * p = NULL;
        memcpy(&hdl, all[j].handle, sizeof(hdl));
        if (cudaIpcOpenMemHandle(&p, hdl, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mine_ok = 0; break; }
        c->peer.box[j] = p;
    }
    char flag = (char) mine_ok, flags[MPS_MAX_RANKS];
    mpsort_comm_allgather_host(c, &flag, flags, 1);
    for (j = 0; j < c->size; j++) if (!flags[j]) ok = 0;
    if (!ok) {
        for (j = 0; j < c->size; j++)
            if (j != c->rank && c->peer.box[j]) { cudaIpcCloseMemHandle(c->peer.box[j]); c->peer.box[j] = NULL; }
        c->peer.state = -1;
        if (c->rank == 0) fprintf(stderr, "MPSort: mailboxes of the peer splitter kernel cannot be mapped; using ncclAllReduce per level\n");
        return 0;
    }
    c->peer.state = 1;
    return 1;
}

int mps_comm_peer_descent(struct mpsort_comm * c, struct mpsk_keyview kv, size_t n, uint32_t nw,
        uint64_t * d_prefix, const uint64_t * d_target, int ns, int level0, int nlevels)
{
    uint32_t * d_err = (uint32_t *) ((char *) c->peer.mine + mpsk_peer_box_bytes());
    /* rank threads that share a GPU: a cudaMalloc / cudaFree of a slower rank (arena growth before
     * its own launch) would wait for the device, i.e. for my spinning kernel, which waits for that
     * rank's kernel. Launch together: nothing allocates between here and the check. */
    if (c->kind == MPS_T_LOCAL) local_barrier(c);
    const int rc = mpsk_splitter_descent_peer(kv, n, nw, d_prefix, d_target, ns, level0, nlevels,
                                              (uint32_t) c->rank, (uint32_t) c->size, c->peer.box, c->peer.seq, d_err, c->stream);
    c->peer.seq += 256;
    return rc;
}

void mps_comm_peer_descent_check(struct mpsort_comm * c)
{
    uint32_t * h = (uint32_t *) mps_host_stage(c, sizeof(uint32_t));
    CUDA_OK(c, cudaMemcpyAsync(h, (char *) c->peer.mine + mpsk_peer_box_bytes(), sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    if (*h) mps_fatal(c, __FILE__, __LINE__, "the peer splitter kernel waited too long for another rank");
}
