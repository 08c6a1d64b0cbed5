/*
 * mpishim.c -- single-host MPI shim over one shared-memory file (TEST
 * INFRASTRUCTURE; see mpi.h). One process per rank, started by mpirun-shim.
 *
 * Transport: every ordered pair (src, dst) owns a single-producer single-consumer
 * ring of message descriptors; message bytes are copied eagerly into the sender's
 * slice of the shared arena and copied out by the receiver (the classic
 * copy-in/copy-out shared-memory MPI transport). Sends therefore never block on the
 * receiver, which makes every collective below a few lines of send/recv.
 * Messages between one pair on one communicator are matched in FIFO order
 * (MPI's non-overtaking rule), so all collectives can share one internal tag.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "mpi.h"
#include "mpishim_layout.h"

static struct shim_hdr * H = NULL;
static int g_rank = 0, g_size = 1;
static char * g_arena = NULL;       /* my slice */
static size_t g_bump = 0;
static uint64_t g_sent = 0;

struct comm_ent { int used; int ctx; int size; int rank; int * world; };
#define MAX_COMMS 256
static struct comm_ent g_comms[MAX_COMMS];

struct req_ent { int used; int is_recv; void * buf; size_t nbytes; int src_world; int tag; int ctx; };
#define MAX_REQS 4096
static struct req_ent g_reqs[MAX_REQS];

#define TAG_COLL (-7777)

static void shim_die(const char * msg)
{
    fprintf(stderr, "mpishim[rank %d]: %s\n", g_rank, msg);
    if (H) H->abort_flag = 1;
    _exit(86);
}

static inline void shim_pause(unsigned * spins)
{
    if (H->abort_flag) _exit(87);
    if (++(*spins) > 200) { sched_yield(); }
    else __builtin_ia32_pause();
}

static struct shim_ring * ring_of(int src, int dst)
{
    return (struct shim_ring *) ((char *) H + H->rings_off) + ((size_t) src * g_size + dst);
}
static struct shim_rankctl * ctl_of(int r)
{
    return (struct shim_rankctl *) ((char *) H + H->ctl_off) + r;
}
static char * arena_of(int r)
{
    return (char *) H + H->arena_off + (size_t) r * H->arena_cap;
}

static size_t type_size(MPI_Datatype t)
{
    if (t & MPISHIM_DERIVED && t > 0) return (size_t) (t & ~MPISHIM_DERIVED);
    switch (t) {
        case MPI_BYTE: case MPI_CHAR: return 1;
        case MPI_INT: return sizeof(int);
        case MPI_LONG: return sizeof(long);
        case MPI_LONG_LONG: return sizeof(long long);
        case MPI_UNSIGNED_LONG: return sizeof(unsigned long);
        case MPI_DOUBLE: return sizeof(double);
        default: break;
    }
    shim_die("unknown datatype");
    return 0;
}

static struct comm_ent * comm_of(MPI_Comm c)
{
    if (c <= 0 || c >= MAX_COMMS || !g_comms[c].used) shim_die("invalid communicator");
    return &g_comms[c];
}

/* ------------------------------------------------------------------------- */
/* point to point on world ranks                                             */

static void raw_send(int dst_world, int ctx, int tag, const void * buf, size_t nbytes)
{
    unsigned spins = 0;
    struct shim_rankctl * me = ctl_of(g_rank);
    const size_t need = (nbytes + 63) & ~(size_t) 63;
    if (need > H->arena_cap) shim_die("message larger than the per-rank arena: raise MPISHIM_ARENA_MB");
    if (__atomic_load_n(&me->consumed, __ATOMIC_ACQUIRE) == g_sent) g_bump = 0;
    while (g_bump + need > H->arena_cap) {
        if (__atomic_load_n(&me->consumed, __ATOMIC_ACQUIRE) == g_sent) { g_bump = 0; break; }
        shim_pause(&spins);
    }
    const size_t off = g_bump;
    if (nbytes) memcpy(g_arena + off, buf, nbytes);
    g_bump += need;

    struct shim_ring * r = ring_of(g_rank, dst_world);
    spins = 0;
    while (r->tail - __atomic_load_n(&r->head, __ATOMIC_ACQUIRE) >= SHIM_RING) shim_pause(&spins);
    struct shim_msg * m = &r->m[r->tail % SHIM_RING];
    m->ctx = ctx; m->tag = tag; m->off = off; m->nbytes = nbytes;
    __atomic_store_n(&m->state, SHIM_FULL, __ATOMIC_RELEASE);
    __atomic_store_n(&r->tail, r->tail + 1, __ATOMIC_RELEASE);
    g_sent++;
}

static void raw_recv(int src_world, int ctx, int tag, void * buf, size_t nbytes)
{
    unsigned spins = 0;
    struct shim_ring * r = ring_of(src_world, g_rank);
    for (;;) {
        const uint64_t head = r->head;
        const uint64_t tail = __atomic_load_n(&r->tail, __ATOMIC_ACQUIRE);
        uint64_t i;
        for (i = head; i < tail; i++) {
            struct shim_msg * m = &r->m[i % SHIM_RING];
            if (__atomic_load_n(&m->state, __ATOMIC_ACQUIRE) != SHIM_FULL) continue;
            if (m->ctx != ctx || m->tag != tag) continue;
            if (m->nbytes > nbytes) shim_die("message truncated: receive buffer too small");
            if (m->nbytes) memcpy(buf, arena_of(src_world) + m->off, m->nbytes);
            m->state = SHIM_DONE;
            __atomic_fetch_add(&ctl_of(src_world)->consumed, 1, __ATOMIC_RELEASE);
            /* retire consumed entries at the head */
            uint64_t h = r->head;
            while (h < tail && r->m[h % SHIM_RING].state == SHIM_DONE) {
                r->m[h % SHIM_RING].state = SHIM_EMPTY;
                h++;
            }
            __atomic_store_n(&r->head, h, __ATOMIC_RELEASE);
            return;
        }
        shim_pause(&spins);
    }
}

static void csend(struct comm_ent * c, int dst, int tag, const void * buf, size_t nbytes)
{
    if (dst < 0 || dst >= c->size) shim_die("send: bad destination rank");
    raw_send(c->world[dst], c->ctx, tag, buf, nbytes);
}
static void crecv(struct comm_ent * c, int src, int tag, void * buf, size_t nbytes)
{
    if (src < 0 || src >= c->size) shim_die("recv: bad source rank");
    raw_recv(c->world[src], c->ctx, tag, buf, nbytes);
}

/* ------------------------------------------------------------------------- */
/* init / finalize                                                           */

int MPI_Init(int * argc, char *** argv)
{
    (void) argc; (void) argv;
    const char * file = getenv("MPISHIM_FILE");
    const char * srank = getenv("MPISHIM_RANK");
    const char * ssize = getenv("MPISHIM_SIZE");
    size_t total;
    if (file && srank && ssize) {
        g_rank = atoi(srank);
        g_size = atoi(ssize);
        int fd = open(file, O_RDWR);
        if (fd < 0) { perror("mpishim: open shared file"); _exit(85); }
        struct stat st;
        fstat(fd, &st);
        total = (size_t) st.st_size;
        H = (struct shim_hdr *) mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, fd, 0);
        close(fd);
        if (H == MAP_FAILED) { perror("mpishim: mmap"); _exit(85); }
        if (H->size != g_size) shim_die("shared file does not match MPISHIM_SIZE");
    } else {
        /* singleton */
        g_rank = 0; g_size = 1;
        total = shim_layout_bytes(1, (size_t) 64 << 20);
        H = (struct shim_hdr *) mmap(NULL, total, PROT_READ | PROT_WRITE,
                                     MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (H == MAP_FAILED) { perror("mpishim: mmap"); _exit(85); }
        shim_layout_init(H, 1, (size_t) 64 << 20);
    }
    g_arena = arena_of(g_rank);
    memset(g_comms, 0, sizeof(g_comms));
    int i;
    g_comms[MPI_COMM_WORLD].used = 1;
    g_comms[MPI_COMM_WORLD].ctx = 1;
    g_comms[MPI_COMM_WORLD].size = g_size;
    g_comms[MPI_COMM_WORLD].rank = g_rank;
    g_comms[MPI_COMM_WORLD].world = (int *) malloc(sizeof(int) * g_size);
    for (i = 0; i < g_size; i++) g_comms[MPI_COMM_WORLD].world[i] = i;
    g_comms[MPI_COMM_SELF].used = 1;
    g_comms[MPI_COMM_SELF].ctx = 2;
    g_comms[MPI_COMM_SELF].size = 1;
    g_comms[MPI_COMM_SELF].rank = 0;
    g_comms[MPI_COMM_SELF].world = (int *) malloc(sizeof(int));
    g_comms[MPI_COMM_SELF].world[0] = g_rank;
    return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
    MPI_Barrier(MPI_COMM_WORLD);
    return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int errorcode)
{
    (void) comm;
    fprintf(stderr, "mpishim[rank %d]: MPI_Abort(%d)\n", g_rank, errorcode);
    fflush(NULL);
    if (H) H->abort_flag = 1;
    _exit((errorcode & 0xff) ? (errorcode & 0xff) : 1);
    return 0;
}

double MPI_Wtime(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

/* ------------------------------------------------------------------------- */
/* communicators                                                             */

int MPI_Comm_size(MPI_Comm comm, int * size) { *size = comm_of(comm)->size; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int * rank) { *rank = comm_of(comm)->rank; return MPI_SUCCESS; }

int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm * newcomm)
{
    struct comm_ent * c = comm_of(comm);
    const int p = c->size;
    int mine[2] = { color, key };
    int * all = (int *) malloc(sizeof(int) * 2 * (size_t) p);
    MPI_Allgather(mine, 2, MPI_INT, all, 2, MPI_INT, comm);
    /* members of my color ordered by (key, parent rank) */
    int * members = (int *) malloc(sizeof(int) * (size_t) p);
    int n = 0, i, j;
    for (i = 0; i < p; i++) if (all[2 * i] == color) members[n++] = i;
    for (i = 1; i < n; i++) {           /* insertion sort, stable */
        const int m = members[i];
        for (j = i; j > 0 && all[2 * members[j - 1] + 1] > all[2 * m + 1]; j--) members[j] = members[j - 1];
        members[j] = m;
    }
    /* the first member allocates a fresh context id and tells the others */
    int ctx = 0;
    if (members[0] == c->rank) {
        ctx = __atomic_fetch_add(&H->next_ctx, 1, __ATOMIC_RELAXED);
        for (i = 1; i < n; i++) csend(c, members[i], TAG_COLL - 1, &ctx, sizeof(ctx));
    } else {
        crecv(c, members[0], TAG_COLL - 1, &ctx, sizeof(ctx));
    }
    if (color == MPI_UNDEFINED) {
        *newcomm = MPI_COMM_NULL;
    } else {
        int slot;
        for (slot = 3; slot < MAX_COMMS && g_comms[slot].used; slot++);
        if (slot == MAX_COMMS) shim_die("too many communicators");
        g_comms[slot].used = 1;
        g_comms[slot].ctx = ctx;
        g_comms[slot].size = n;
        g_comms[slot].world = (int *) malloc(sizeof(int) * (size_t) n);
        for (i = 0; i < n; i++) {
            g_comms[slot].world[i] = c->world[members[i]];
            if (members[i] == c->rank) g_comms[slot].rank = i;
        }
        *newcomm = slot;
    }
    free(members);
    free(all);
    return MPI_SUCCESS;
}

int MPI_Comm_free(MPI_Comm * comm)
{
    struct comm_ent * c = comm_of(*comm);
    if (*comm > MPI_COMM_SELF) {
        free(c->world);
        c->used = 0;
    }
    *comm = MPI_COMM_NULL;
    return MPI_SUCCESS;
}

/* ------------------------------------------------------------------------- */
/* datatypes                                                                 */

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype * newtype)
{
    const size_t n = (size_t) count * type_size(oldtype);
    if (n >= (size_t) MPISHIM_DERIVED) shim_die("derived type too large");
    *newtype = (MPI_Datatype) (MPISHIM_DERIVED | (int) n);
    return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype * type) { (void) type; return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype * type) { *type = MPI_DATATYPE_NULL; return MPI_SUCCESS; }
int MPI_Type_get_extent(MPI_Datatype type, MPI_Aint * lb, MPI_Aint * extent)
{
    if (lb) *lb = 0;
    *extent = (MPI_Aint) type_size(type);
    return MPI_SUCCESS;
}

/* ------------------------------------------------------------------------- */
/* collectives (linear algorithms; p is small)                               */

int MPI_Barrier(MPI_Comm comm)
{
    struct comm_ent * c = comm_of(comm);
    int i;
    if (c->size == 1) return MPI_SUCCESS;
    if (c->rank == 0) {
        for (i = 1; i < c->size; i++) crecv(c, i, TAG_COLL, NULL, 0);
        for (i = 1; i < c->size; i++) csend(c, i, TAG_COLL, NULL, 0);
    } else {
        csend(c, 0, TAG_COLL, NULL, 0);
        crecv(c, 0, TAG_COLL, NULL, 0);
    }
    return MPI_SUCCESS;
}

int MPI_Bcast(void * buf, int count, MPI_Datatype type, int root, MPI_Comm comm)
{
    struct comm_ent * c = comm_of(comm);
    const size_t n = (size_t) count * type_size(type);
    int i;
    if (c->rank == root) {
        for (i = 0; i < c->size; i++) if (i != root) csend(c, i, TAG_COLL, buf, n);
    } else {
        crecv(c, root, TAG_COLL, buf, n);
    }
    return MPI_SUCCESS;
}

static void reduce_into(void * acc, const void * in, int count, MPI_Datatype type, MPI_Op op)
{
    int i;
#define RED(T, UT) do { T * a = (T *) acc; const T * b = (const T *) in; \
        for (i = 0; i < count; i++) { \
            if (op == MPI_SUM) a[i] = (T) ((UT) a[i] + (UT) b[i]); \
            else if (op == MPI_MIN) { if (b[i] < a[i]) a[i] = b[i]; } \
            else if (op == MPI_MAX) { if (b[i] > a[i]) a[i] = b[i]; } \
            else shim_die("unsupported reduction op"); } } while (0)
    switch (type) {
        case MPI_INT: RED(int, unsigned int); break;
        case MPI_LONG: RED(long, unsigned long); break;
        case MPI_LONG_LONG: RED(long long, unsigned long long); break;
        case MPI_UNSIGNED_LONG: RED(unsigned long, unsigned long); break;
        case MPI_DOUBLE: RED(double, double); break;
        default: shim_die("unsupported reduction datatype");
    }
#undef RED
}

int MPI_Allreduce(const void * sendbuf, void * recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm)
{
    struct comm_ent * c = comm_of(comm);
    const size_t n = (size_t) count * type_size(type);
    int i;
    if (sendbuf != MPI_IN_PLACE && sendbuf != recvbuf) memcpy(recvbuf, sendbuf, n);
    if (c->size == 1) return MPI_SUCCESS;
    if (c->rank == 0) {
        void * tmp = malloc(n ? n : 1);
        for (i = 1; i < c->size; i++) {       /* rank order: deterministic */
            crecv(c, i, TAG_COLL, tmp, n);
            reduce_into(recvbuf, tmp, count, type, op);
        }
        free(tmp);
        for (i = 1; i < c->size; i++) csend(c, i, TAG_COLL, recvbuf, n);
    } else {
        csend(c, 0, TAG_COLL, recvbuf, n);
        crecv(c, 0, TAG_COLL, recvbuf, n);
    }
    return MPI_SUCCESS;
}

int MPI_Gatherv(const void * sendbuf, int sendcount, MPI_Datatype sendtype,
                void * recvbuf, const int * recvcounts, const int * displs, MPI_Datatype recvtype,
                int root, MPI_Comm comm)
{
    struct comm_ent * c = comm_of(comm);
    int i;
    if (c->rank == root) {
        const size_t rs = type_size(recvtype);
        for (i = 0; i < c->size; i++) {
            char * dst = (char *) recvbuf + (size_t) displs[i] * rs;
            const size_t n = (size_t) recvcounts[i] * rs;
            if (i == root) {
                if (sendbuf != MPI_IN_PLACE && n) memmove(dst, sendbuf, n);
            } else {
                crecv(c, i, TAG_COLL, dst, n);
            }
        }
    } else {
        csend(c, root, TAG_COLL, sendbuf, (size_t) sendcount * type_size(sendtype));
    }
    return MPI_SUCCESS;
}

int MPI_Gather(const void * sendbuf, int sendcount, MPI_Datatype sendtype,
               void * recvbuf, int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm)
{
    struct comm_ent * c = comm_of(comm);
    int i, rc;
    int * counts = (int *) malloc(sizeof(int) * 2 * (size_t) c->size);
    int * displs = counts + c->size;
    for (i = 0; i < c->size; i++) { counts[i] = recvcount; displs[i] = i * recvcount; }
    rc = MPI_Gatherv(sendbuf, sendcount, sendtype, recvbuf, counts, displs, recvtype, root, comm);
    free(counts);
    return rc;
}

int MPI_Scatterv(const void * sendbuf, const int * sendcounts, const int * displs, MPI_Datatype sendtype,
                 void * recvbuf, int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm)
{
    struct comm_ent * c = comm_of(comm);
    int i;
    if (c->rank == root) {
        const size_t ss = type_size(sendtype);
        for (i = 0; i < c->size; i++) {
            const char * src = (const char *) sendbuf + (size_t) displs[i] * ss;
            const size_t n = (size_t) sendcounts[i] * ss;
            if (i == root) {
                if (recvbuf != MPI_IN_PLACE && n) memmove(recvbuf, src, n);
            } else {
                csend(c, i, TAG_COLL, src, n);
            }
        }
    } else {
        crecv(c, root, TAG_COLL, recvbuf, (size_t) recvcount * type_size(recvtype));
    }
    return MPI_SUCCESS;
}

int MPI_Allgather(const void * sendbuf, int sendcount, MPI_Datatype sendtype,
                  void * recvbuf, int recvcount, MPI_Datatype recvtype, MPI_Comm comm)
{
    struct comm_ent * c = comm_of(comm);
    const size_t n = (size_t) recvcount * type_size(recvtype);
    int i;
    (void) sendcount; (void) sendtype;
    /* MPI_IN_PLACE: my piece already sits at recvbuf + rank*n (mp-mpiu.c:416) */
    const char * mine = (sendbuf == MPI_IN_PLACE) ? (const char *) recvbuf + (size_t) c->rank * n
                                                   : (const char *) sendbuf;
    if (sendbuf != MPI_IN_PLACE && n) memmove((char *) recvbuf + (size_t) c->rank * n, mine, n);
    for (i = 0; i < c->size; i++) if (i != c->rank) csend(c, i, TAG_COLL, mine, n);
    for (i = 0; i < c->size; i++) if (i != c->rank) crecv(c, i, TAG_COLL, (char *) recvbuf + (size_t) i * n, n);
    return MPI_SUCCESS;
}

int MPI_Alltoallv(const void * sendbuf, const int * sendcounts, const int * sdispls, MPI_Datatype sendtype,
                  void * recvbuf, const int * recvcounts, const int * rdispls, MPI_Datatype recvtype, MPI_Comm comm)
{
    struct comm_ent * c = comm_of(comm);
    const size_t ss = type_size(sendtype), rs = type_size(recvtype);
    int k;
    /* shifted order spreads the load over the peers */
    for (k = 1; k < c->size; k++) {
        const int dst = (c->rank + k) % c->size;
        csend(c, dst, TAG_COLL, (const char *) sendbuf + (size_t) sdispls[dst] * ss, (size_t) sendcounts[dst] * ss);
    }
    if (sendcounts[c->rank])
        memmove((char *) recvbuf + (size_t) rdispls[c->rank] * rs,
                (const char *) sendbuf + (size_t) sdispls[c->rank] * ss, (size_t) sendcounts[c->rank] * ss);
    for (k = 1; k < c->size; k++) {
        const int src = (c->rank - k + c->size) % c->size;
        crecv(c, src, TAG_COLL, (char *) recvbuf + (size_t) rdispls[src] * rs, (size_t) recvcounts[src] * rs);
    }
    return MPI_SUCCESS;
}

int MPI_Alltoall(const void * sendbuf, int sendcount, MPI_Datatype sendtype,
                 void * recvbuf, int recvcount, MPI_Datatype recvtype, MPI_Comm comm)
{
    struct comm_ent * c = comm_of(comm);
    int i, rc;
    int * v = (int *) malloc(sizeof(int) * 4 * (size_t) c->size);
    for (i = 0; i < c->size; i++) {
        v[i] = sendcount; v[c->size + i] = i * sendcount;
        v[2 * c->size + i] = recvcount; v[3 * c->size + i] = i * recvcount;
    }
    rc = MPI_Alltoallv(sendbuf, v, v + c->size, sendtype, recvbuf, v + 2 * c->size, v + 3 * c->size, recvtype, comm);
    free(v);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* user point to point                                                       */

int MPI_Send(const void * buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm)
{
    csend(comm_of(comm), dest, tag, buf, (size_t) count * type_size(type));
    return MPI_SUCCESS;
}

int MPI_Recv(void * buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status * status)
{
    crecv(comm_of(comm), source, tag, buf, (size_t) count * type_size(type));
    if (status) { status->MPI_SOURCE = source; status->MPI_TAG = tag; status->MPI_ERROR = MPI_SUCCESS; }
    return MPI_SUCCESS;
}

int MPI_Sendrecv(const void * sendbuf, int sendcount, MPI_Datatype sendtype, int dest, int sendtag,
                 void * recvbuf, int recvcount, MPI_Datatype recvtype, int source, int recvtag,
                 MPI_Comm comm, MPI_Status * status)
{
    MPI_Send(sendbuf, sendcount, sendtype, dest, sendtag, comm);
    return MPI_Recv(recvbuf, recvcount, recvtype, source, recvtag, comm, status);
}

int MPI_Isend(const void * buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request * req)
{
    /* eager: complete on return */
    MPI_Send(buf, count, type, dest, tag, comm);
    *req = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}

int MPI_Irecv(void * buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request * req)
{
    struct comm_ent * c = comm_of(comm);
    int slot;
    for (slot = 0; slot < MAX_REQS && g_reqs[slot].used; slot++);
    if (slot == MAX_REQS) shim_die("too many outstanding requests");
    if (source < 0 || source >= c->size) shim_die("irecv: bad source rank");
    g_reqs[slot].used = 1;
    g_reqs[slot].is_recv = 1;
    g_reqs[slot].buf = buf;
    g_reqs[slot].nbytes = (size_t) count * type_size(type);
    g_reqs[slot].src_world = c->world[source];
    g_reqs[slot].tag = tag;
    g_reqs[slot].ctx = c->ctx;
    *req = slot;
    return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request * req, MPI_Status * status)
{
    (void) status;
    if (*req == MPI_REQUEST_NULL) return MPI_SUCCESS;
    struct req_ent * r = &g_reqs[*req];
    if (!r->used) shim_die("wait on an inactive request");
    raw_recv(r->src_world, r->ctx, r->tag, r->buf, r->nbytes);
    r->used = 0;
    *req = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}

int MPI_Waitall(int count, MPI_Request * reqs, MPI_Status * statuses)
{
    int i;
    (void) statuses;
    for (i = 0; i < count; i++) MPI_Wait(&reqs[i], MPI_STATUS_IGNORE);
    return MPI_SUCCESS;
}
