/*
 * mpi_facade.c -- the few MPI calls the reference's C drivers make for themselves, over the C ABI
 * of libmpsort-b200.so (TEST / INTEGRATION GLUE; see mpi.h). It moves bytes and sums integers;
 * every sort goes through the product's mpsort_mpi_impl / mpsort_mpi_newarray_impl unchanged.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "mpi.h"

/* the product's C ABI (include/mpsort.h); declared here instead of including that header because the
 * driver's translation unit sees the REFERENCE's mpsort.h and this file must agree with both */
#define MPSORT_UNIQUE_ID_BYTES 128
int mpsort_comm_get_unique_id(void * id);
MPI_Comm mpsort_comm_init_rank(int rank, int size, const void * unique_id, int device);
MPI_Comm mpsort_comm_self(int device);
int mpsort_comm_init_local_group(int size, const int * devices, MPI_Comm * comms);
void mpsort_comm_destroy(MPI_Comm comm);
int mpsort_comm_rank(MPI_Comm comm);
int mpsort_comm_size(MPI_Comm comm);
void mpsort_comm_barrier(MPI_Comm comm);
void mpsort_comm_allgather_host(MPI_Comm comm, const void * send, void * recv, size_t nbytes_per_rank);

static __thread MPI_Comm t_world;
static __thread unsigned t_seq_send[64], t_seq_recv[64];       /* per peer: messages are matched in order */
static char g_job[200];

MPI_Comm mpsort_dropin_world(void)
{
    if (!t_world) { fprintf(stderr, "dropin: MPI_COMM_WORLD used before MPI_Init\n"); abort(); }
    return t_world;
}

static size_t type_size(MPI_Datatype t)
{
    switch (t) {
    case MPI_BYTE: return 1;
    case MPI_INT: return 4;
    case MPI_LONG: case MPI_LONG_LONG: case MPI_DOUBLE: return 8;
    default: fprintf(stderr, "dropin: datatype %d not supported\n", t); abort();
    }
}

static void job_name(void)
{
    const char * e = getenv("MPSORT_DROPIN_JOB");
    if (e) snprintf(g_job, sizeof(g_job), "%s", e);
    else snprintf(g_job, sizeof(g_job), "/dev/shm/mpsort_dropin_%d", (int) getpid());
}

static void write_file(const char * path, const void * buf, size_t n)
{
    char tmp[300];
    snprintf(tmp, sizeof(tmp), "%s.tmp", path);
    int fd = open(tmp, O_CREAT | O_WRONLY | O_TRUNC, 0600);
    if (fd < 0 || (n && write(fd, buf, n) != (ssize_t) n)) { perror("dropin: write"); abort(); }
    close(fd);
    if (rename(tmp, path) != 0) { perror("dropin: rename"); abort(); }
}

static void read_file(const char * path, void * buf, size_t n, int unlink_after)
{
    const double deadline = MPI_Wtime() + 300.0;
    int fd;
    while ((fd = open(path, O_RDONLY)) < 0) {
        if (MPI_Wtime() > deadline) { fprintf(stderr, "dropin: %s did not appear\n", path); abort(); }
        usleep(200);
    }
    if (n && read(fd, buf, n) != (ssize_t) n) { fprintf(stderr, "dropin: short message in %s\n", path); abort(); }
    close(fd);
    if (unlink_after) unlink(path);
}

int MPI_Init(int * argc, char *** argv)
{
    (void) argc; (void) argv;
    if (t_world) return MPI_SUCCESS;                 /* thread mode: the facade's main() made it */
    const int rank = atoi(getenv("RANK") ? getenv("RANK") : "0");
    const int size = atoi(getenv("WORLD_SIZE") ? getenv("WORLD_SIZE") : "1");
    const int local = atoi(getenv("LOCAL_RANK") ? getenv("LOCAL_RANK") : "0");
    job_name();
    if (size <= 1) { t_world = mpsort_comm_self(local); return MPI_SUCCESS; }
    if (!getenv("MPSORT_DROPIN_JOB")) { fprintf(stderr, "dropin: WORLD_SIZE > 1 needs MPSORT_DROPIN_JOB (use tests/dropin/launch.py)\n"); abort(); }
    /* the NCCL id travels through a file, as an application with a real MPI would MPI_Bcast it */
    char id[MPSORT_UNIQUE_ID_BYTES], path[300];
    snprintf(path, sizeof(path), "%s.ncclid", g_job);
    if (rank == 0) {
        if (mpsort_comm_get_unique_id(id) != 0) abort();
        write_file(path, id, sizeof(id));
    } else read_file(path, id, sizeof(id), 0);
    t_world = mpsort_comm_init_rank(rank, size, id, local);
    mpsort_comm_barrier(t_world);
    if (rank == 0) unlink(path);
    return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
    if (t_world) { mpsort_comm_barrier(t_world); mpsort_comm_destroy(t_world); t_world = NULL; }
    return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int code) { (void) comm; fprintf(stderr, "dropin: MPI_Abort(%d)\n", code); fflush(NULL); _exit(code ? code & 255 ? code & 255 : 1 : 1); }
int MPI_Comm_rank(MPI_Comm comm, int * rank) { *rank = mpsort_comm_rank(comm); return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int * size) { *size = mpsort_comm_size(comm); return MPI_SUCCESS; }

static void rank_reached_a_collective(void);

int MPI_Barrier(MPI_Comm comm) { rank_reached_a_collective(); mpsort_comm_barrier(comm); return MPI_SUCCESS; }

double MPI_Wtime(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double) t.tv_sec + 1e-9 * (double) t.tv_nsec;
}

int MPI_Allreduce(const void * send, void * recv, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm)
{
    const int p = mpsort_comm_size(comm);
    const size_t w = type_size(type), nbytes = w * (size_t) count;
    char * all = (char *) malloc(nbytes * (size_t) p + 1);
    int i, r;
    rank_reached_a_collective();
    mpsort_comm_allgather_host(comm, send == MPI_IN_PLACE ? recv : send, all, nbytes);
    for (i = 0; i < count; i++) {
        if (w == 4) {
            int acc = ((const int *) all)[i];
            for (r = 1; r < p; r++) {
                const int v = ((const int *) (all + nbytes * r))[i];
                acc = op == MPI_SUM ? acc + v : op == MPI_MIN ? (v < acc ? v : acc) : (v > acc ? v : acc);
            }
            ((int *) recv)[i] = acc;
        } else if (type == MPI_DOUBLE) {
            double acc = ((const double *) all)[i];
            for (r = 1; r < p; r++) {
                const double v = ((const double *) (all + nbytes * r))[i];
                acc = op == MPI_SUM ? acc + v : op == MPI_MIN ? (v < acc ? v : acc) : (v > acc ? v : acc);
            }
            ((double *) recv)[i] = acc;
        } else {
            long long acc = ((const long long *) all)[i];
            for (r = 1; r < p; r++) {
                const long long v = ((const long long *) (all + nbytes * r))[i];
                acc = op == MPI_SUM ? (long long) ((unsigned long long) acc + (unsigned long long) v)
                    : op == MPI_MIN ? (v < acc ? v : acc) : (v > acc ? v : acc);
            }
            ((long long *) recv)[i] = acc;
        }
    }
    free(all);
    return MPI_SUCCESS;
}

int MPI_Bcast(void * buf, int count, MPI_Datatype type, int root, MPI_Comm comm)
{
    const int p = mpsort_comm_size(comm);
    const size_t nbytes = type_size(type) * (size_t) count;
    char * all = (char *) malloc(nbytes * (size_t) p + 1);
    rank_reached_a_collective();
    mpsort_comm_allgather_host(comm, buf, all, nbytes);
    memcpy(buf, all + nbytes * (size_t) root, nbytes);
    free(all);
    return MPI_SUCCESS;
}

/* point to point (the drivers' neighbour checks, bench-mpi.c:66-96, main-mpi.c:67-93): one small file per message */
static void msg_path(char * path, size_t n, int src, int dst, int tag, unsigned seq)
{
    snprintf(path, n, "%s.msg.%d.%d.%d.%u", g_job, src, dst, tag, seq);
}

int MPI_Send(const void * buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm)
{
    char path[300];
    msg_path(path, sizeof(path), mpsort_comm_rank(comm), dest, tag, t_seq_send[dest & 63]++);
    write_file(path, buf, type_size(type) * (size_t) count);
    return MPI_SUCCESS;
}

int MPI_Recv(void * buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm, MPI_Status * st)
{
    char path[300];
    (void) st;
    msg_path(path, sizeof(path), src, mpsort_comm_rank(comm), tag, t_seq_recv[src & 63]++);
    read_file(path, buf, type_size(type) * (size_t) count, 1);
    return MPI_SUCCESS;
}

int MPI_Sendrecv(const void * sbuf, int scount, MPI_Datatype stype, int dest, int stag,
                 void * rbuf, int rcount, MPI_Datatype rtype, int src, int rtag, MPI_Comm comm, MPI_Status * st)
{
    MPI_Send(sbuf, scount, stype, dest, stag, comm);          /* sends never block here */
    return MPI_Recv(rbuf, rcount, rtype, src, rtag, comm, st);
}

/* ------------------------------------------------------------------------- */
/* thread mode (MPSORT_DROPIN_THREADS=P): built with -Dmain=mpsort_dropin_main  */
#ifdef MPSORT_DROPIN_OWNS_MAIN
#undef main
static int g_argc;
static char ** g_argv;
static MPI_Comm g_comms[64];
static int g_rc[64];
static pthread_mutex_t g_start_lock = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_start_cond = PTHREAD_COND_INITIALIZER;
static int g_started_next;           /* ranks 0 .. g_started_next-1 have reached their first collective */
static __thread int t_rank_index = -1, t_signalled;

/* getopt's cursor is process-global (bench-mpi.c:130 parses its flags after MPI_Init): rank k+1 starts
 * only when rank k has reached its first collective, i.e. is past its argument parsing */
static void rank_reached_a_collective(void)
{
    if (t_rank_index < 0 || t_signalled) return;
    t_signalled = 1;
    pthread_mutex_lock(&g_start_lock);
    g_started_next = t_rank_index + 1;
    pthread_cond_broadcast(&g_start_cond);
    pthread_mutex_unlock(&g_start_lock);
}

static void * rank_thread(void * arg)
{
    const int r = (int) (intptr_t) arg;
    extern int optind;
    pthread_mutex_lock(&g_start_lock);
    while (g_started_next < r) pthread_cond_wait(&g_start_cond, &g_start_lock);
    optind = 1;
    pthread_mutex_unlock(&g_start_lock);
    t_rank_index = r;
    t_world = g_comms[r];
    g_rc[r] = mpsort_dropin_main(g_argc, g_argv);
    rank_reached_a_collective();                      /* a rank that left early must not hold up the others */
    return NULL;
}

int main(int argc, char ** argv)
{
    const char * e = getenv("MPSORT_DROPIN_THREADS");
    const int p = e ? atoi(e) : 0;
    int r, rc = 0, devices[64] = { 0 };
    pthread_t th[64];
    if (p <= 1) return mpsort_dropin_main(argc, argv);        /* process mode or singleton: MPI_Init decides */
    if (p > 64) { fprintf(stderr, "dropin: at most 64 rank threads\n"); return 2; }
    job_name();
    g_argc = argc; g_argv = argv;
    if (mpsort_comm_init_local_group(p, devices, g_comms) != 0) { fprintf(stderr, "dropin: local group failed\n"); return 2; }
    for (r = 0; r < p; r++) pthread_create(&th[r], NULL, rank_thread, (void *) (intptr_t) r);
    for (r = 0; r < p; r++) { pthread_join(th[r], NULL); rc |= g_rc[r]; }
    return rc;
}
#else
static void rank_reached_a_collective(void) { }
#endif
