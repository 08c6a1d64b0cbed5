/*
 * mpsort_internal.h -- shared declarations of the C host code of mpsort-b200.
 */
#ifndef MPSORT_INTERNAL_H
#define MPSORT_INTERNAL_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <pthread.h>

#include <cuda_runtime_api.h>
#include <nccl.h>

#include "mpsort.h"
#include "mpsort_kernels.h"

#define MPS_MAX_RANKS 64
#define MPS_MAX_KEY_WORDS 16

enum mps_transport { MPS_T_SELF = 0, MPS_T_NCCL = 1, MPS_T_LOCAL = 2 };

/* device arena slots (grow-only, per communicator) */
enum mps_slot {
    MPS_S_DIN = 0,     /* staged copy of a host input                  */
    MPS_S_DOUT,        /* staged copy of a host output                 */
    MPS_S_KW,          /* key words in original order  [nw][n] u64     */
    MPS_S_KB,          /* ping-pong key buffer         [n] u64         */
    MPS_S_KA,          /* second ping-pong key buffer (multi-word)     */
    MPS_S_IA,          /* ping-pong index buffers      [n] u32         */
    MPS_S_IB,
    MPS_S_SK,          /* sorted key words (multi-word) [nw][n] u64    */
    MPS_S_HIST,        /* histograms + scanned bins                    */
    MPS_S_SCRATCH,     /* look-back status words                       */
    MPS_S_SEND,        /* packed records, destination-contiguous       */
    MPS_S_RECV,        /* received runs, source-rank order             */
    MPS_S_SPLIT,       /* splitter state: prefix, target, counts, out  */
    MPS_S_STAGE,       /* device side of small host collectives        */
    MPS_S_STAGE2,      /* scratch for the in-process allreduce         */
    MPS_S_MISC,        /* checksum / small outputs                     */
    MPS_S_MERGE_SAMP,  /* merge: sample keys                           */
    MPS_S_MERGE_CUT,   /* merge: per-tile cut positions                */
    MPS_S_MERGE_SORTED,/* merge: sample keys in merged order            */
    MPS_S_MERGE_SID,   /* merge: their sample ids                       */
    MPS_S_PRED,        /* hybrid predictor: hash tables + pair counts   */
    MPS_NSLOTS
};

struct mps_local_group {
    int size;
    pthread_barrier_t barrier;
    pthread_mutex_t lock;
    int refcount;
    const void * slot[MPS_MAX_RANKS];
    const void * slot2[MPS_MAX_RANKS];
    size_t val[MPS_MAX_RANKS];
};

#define MPS_MAX_TIMERS 48
struct mps_timers {
    cudaEvent_t ev[MPS_MAX_TIMERS];
    char name[MPS_MAX_TIMERS][20];
    int n;
    int created;
};

/* per-kernel-class timing (bench.py's roofline numbers): CUDA events around every
 * launch of a class on the communicator's stream, summed after the stream sync */
enum mps_kclass { MPS_K_EXTRACT = 0, MPS_K_ONESWEEP, MPS_K_ONESWEEP_REC, MPS_K_GATHER_KEYS, MPS_K_GATHER_RECORDS,
                  MPS_K_SPLITTER, MPS_K_CHECKSUM, MPS_K_EXCHANGE, MPS_K_MERGE, MPS_K_HYBRID, MPS_NKCLASS };
#define MPS_KT_MAX 512
struct mps_ktimes {
    int on;
    int force_cls;                       /* >= 0: book every launch under this class */
    int n;                               /* pending event pairs */
    int nev;                             /* events created */
    cudaEvent_t ev[2 * MPS_KT_MAX];
    int cls[MPS_KT_MAX];
    double ms[MPS_NKCLASS];
    uint64_t launches[MPS_NKCLASS];
};

struct mpsort_comm {
    int kind;
    int rank, size, device;
    cudaStream_t stream;
    ncclComm_t nccl;
    struct mps_local_group * grp;

    struct { void * ptr; size_t cap; } slot[MPS_NSLOTS];
    void * h_stage;        /* pinned */
    size_t h_stage_cap;

    int verbose_malloc;
    struct mps_timers timers;

    struct mpsort_last_stats stats;
    int64_t sendcounts[MPS_MAX_RANKS];
    struct mps_ktimes kt;
    cudaStream_t stream2;                      /* merges of the pipelined exchange run here */
    /* merge tiles that broke their bound (never happens) are counted on the device and looked at
     * once per sort, after the final synchronisation, instead of once per exchange part */
    uint32_t * d_merge_ovf;                    /* device u32: tiles that broke their bound */
    uint32_t * h_merge_ovf;                    /* pinned copy */
    int merge_ovf_pending;

    /* Host buffers. A large host INPUT arrives in chunks on its own copy stream and the
     * histogram / key-extraction pass of FirstSort runs chunk by chunk behind it; a host OUTPUT leaves range by
     * range on a second copy stream as soon as a prefix of it is final (fix-up chunks, merged exchange parts),
     * while the rest is still being produced. */
#define MPS_MAX_CHUNKS 32
    struct {
        const void * dbase;                    /* the staged device copy the chunks land in (NULL: not chunked) */
        size_t n, elsize, chunk;               /* records in all, bytes per record, records per chunk */
        int nchunks, waited;                   /* chunks, and how many of them c->stream has waited for */
        cudaEvent_t ev[MPS_MAX_CHUNKS];
    } in;
    struct {
        void * host;                           /* NULL: the output is not a staged host buffer */
        const void * dout;
        size_t elsize, total, done;            /* records: in all / already on their way */
    } outp;
    cudaStream_t h2d_stream, d2h_stream;
    cudaEvent_t io_ev[2];
    int io_created;
    cudaEvent_t phase_ev[MPS_MAX_RANKS + 1];
    int phase_ev_created;

    /* CANDIDATE (MPSORT_PEER_SPLITTER=1): mailboxes of the one-kernel splitter descent */
    struct {
        int state;                             /* 0 not tried yet, 1 ready, -1 unavailable */
        void * mine;                           /* my mailbox + error word (cudaMalloc) */
        void * box[MPS_MAX_RANKS];             /* rank j's mailbox as mapped here */
        uint32_t seq;                          /* 256 * sorts that used the kernel so far */
    } peer;

    /* peer-store exchange (NCCL transport only): every rank's receive buffer mapped here */
    struct {
        int disabled;                          /* env MPSORT_NO_P2P, or a mapping failed somewhere */
        int pull;                              /* 1: read peers' send buffers, 0: write peers' receive buffers */
        void * peer_base[MPS_MAX_RANKS];       /* my mapping of rank j's exchange buffer */
        uint64_t peer_ptr[MPS_MAX_RANKS];      /* rank j's own pointer the mapping belongs to */
        void ** zombies;                       /* my old receive buffers peers may still have mapped (never freed
                                                * before the communicator is destroyed; the list grows as needed) */
        int nzombies, zombies_cap;
        uint64_t generation;                   /* bumped every time my exchange buffer is replaced */
        uint64_t peer_gen[MPS_MAX_RANKS];      /* generation of rank j's buffer my mapping belongs to */
        int * d_flag;                          /* device word for the completion all-reduce */
        int skip_barrier;                      /* 1: the caller asks for the completion barrier itself (after its last part) */
        int burst;                             /* 1: a sparse exchange -- small slices, all copy streams at once */
        int skip_self;                         /* 1: the caller merges its own slice from the send buffer -- no transport copies it (any transport) */
        int chained;                           /* a later part of ONE exchange: what its copies wait for (exchange_p2p: 0 the barrier before, 1 nothing, 2 the barrier before that) */
        int part;                              /* index of the part within its exchange */
        int copy_engine;                       /* >= 1: slices move by cudaMemcpyAsync (DMA engines, no SMs), that many at a time */
        cudaStream_t ce_stream[8];             /* copy-engine mode: peer copies fan out over these */
        cudaEvent_t ce_ev[10];                 /* [0..7] end of a copy stream's part; [8], [9] gates (exchange_p2p) */
        int ce_created;
    } p2p;
};

/* what every rank tells the others about its receive buffer during LayDistr */
struct mps_recv_info { uint64_t ptr; uint64_t cap; uint64_t generation; unsigned char handle[64]; };
void mps_comm_recv_info(struct mpsort_comm * c, void * recvbuf, void * sendbuf, struct mps_recv_info * info);
/* after the all-gather: (re)map peers whose buffer changed; returns 1 if the peer-store
 * exchange can be used by ALL ranks this call (collective decision) */
int mps_comm_p2p_prepare(struct mpsort_comm * c, const struct mps_recv_info * all);

void mps_kt_begin(struct mpsort_comm * c, int cls);
void mps_kt_end(struct mpsort_comm * c);
void mps_kt_collect(struct mpsort_comm * c);   /* call after the stream was synchronised */
#define KERN_T(c, cls, call) do { mps_kt_begin((c), (cls)); KERN_OK((c), call); mps_kt_end((c)); } while (0)

/* ---- errors (reference convention: message incl. caller site, then abort) ---- */
void mps_fatal(struct mpsort_comm * c, const char * file, int line, const char * fmt, ...)
    __attribute__((noreturn, format(printf, 4, 5)));

extern __thread const char * mps_caller_file;
extern __thread int mps_caller_line;

#define CUDA_OK(c, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) \
    mps_fatal((c), __FILE__, __LINE__, "CUDA error %d (%s) in %s", (int) e__, cudaGetErrorString(e__), #call); } while (0)
#define KERN_OK(c, call) do { int e__ = (call); if (e__ != 0) \
    mps_fatal((c), __FILE__, __LINE__, "kernel launch error %d (%s) in %s", e__, cudaGetErrorString((cudaError_t) e__), #call); } while (0)
#define NCCL_OK(c, call) do { ncclResult_t e__ = (call); if (e__ != ncclSuccess) \
    mps_fatal((c), __FILE__, __LINE__, "NCCL error %d (%s) in %s", (int) e__, ncclGetErrorString(e__), #call); } while (0)

/* ---- host allocations of significance go through the mpiu_set_malloc hook ---- */
void * mps_host_malloc(const char * name, size_t size, const char * file, int line);
void mps_host_free(void * ptr, const char * file, int line);

/* merge overflow words: zeroed at the start of a sort, checked after its final synchronisation */
void mps_merge_ovf_begin(struct mpsort_comm * c);
void mps_merge_ovf_fetch(struct mpsort_comm * c);     /* async copy to the pinned words on c->stream */
void mps_merge_ovf_check(struct mpsort_comm * c);     /* after the stream was synchronised */

/* ---- chunked host input / output (mpsort_host.c) ---- */
/* c->stream goes on only after records [first, first + count) of the staged input dbase have arrived
 * (no-op for any other pointer) */
void mps_input_wait(struct mpsort_comm * c, const void * dbase, size_t first, size_t count);
/* the ranges to process dbase in: a[0] = 0 < a[1] < ... < a[nr] = n; returns nr (1 unless dbase is chunked) */
int mps_input_ranges(struct mpsort_comm * c, const void * dbase, size_t n, size_t * a);
/* records [0, upto) of the staged output are final once what is queued on c->stream has run: send on what
 * has not left yet (no-op unless the output is a staged host buffer) */
void mps_output_ready(struct mpsort_comm * c, size_t upto);
/* forget what was announced: everything is sent again at the end (a late change to records already sent) */
void mps_output_reset(struct mpsort_comm * c);

/* ---- arena ---- */
void * mps_arena_get(struct mpsort_comm * c, int slot, size_t bytes);
void * mps_host_stage(struct mpsort_comm * c, size_t bytes);

/* ---- communicator primitives (mpsort_comm.c) ---- */
void mps_comm_allreduce_u64_dev(struct mpsort_comm * c, uint64_t * dptr, size_t count);
/* items of elsize bytes; cut[j*(p+1) + k] = first item of rank j's sorted array that
 * goes to rank k (full matrix, known on every rank) */
void mps_comm_alltoallv_dev(struct mpsort_comm * c, const void * sendbuf, void * recvbuf,
        const int64_t * cut, size_t elsize, int dense, uint64_t * bytes_remote);
/* One phase of the exchange, in items of elsize bytes. I send sendcnt[k] items from item
 * sendoff[k] of sendbuf to rank k; I receive recvcnt[j] items from rank j, stored in
 * source-rank order from item recvoff of recvbuf. peer_recvoff[k] = item of rank k's
 * receive buffer where my slice lands (peer stores); peer_sendoff[j] = item of rank j's
 * send buffer where its slice for me starts (pull transports). use_p2p needs a successful
 * mps_comm_p2p_prepare. Adds the bytes that left this GPU to *bytes_remote. */
void mps_comm_exchange(struct mpsort_comm * c, const void * sendbuf, const int64_t * sendoff, const int64_t * sendcnt,
        void * recvbuf, int64_t recvoff, const int64_t * recvcnt,
        const int64_t * peer_recvoff, const int64_t * peer_sendoff,
        size_t elsize, int dense, int use_p2p, uint64_t * bytes_remote);

/* CANDIDATE (MPSORT_FUSED_PACK=1): the same phase from the unsorted records and the sorted
 * permutation, gather and peer stores in one kernel (push transport, after mps_comm_p2p_prepare) */
void mps_comm_exchange_gather(struct mpsort_comm * c, const void * base, const uint32_t * idx,
        const int64_t * sendoff, const int64_t * sendcnt, void * recvbuf, const int64_t * peer_recvoff,
        size_t elsize, uint64_t * bytes_remote);

/* CANDIDATE (MPSORT_PEER_SPLITTER=1): collective; 1 when every rank has every mailbox mapped */
int mps_comm_peer_boxes_prepare(struct mpsort_comm * c);
/* all descent levels in one kernel (mpsk_splitter_descent_peer); returns a CUDA error code */
int mps_comm_peer_descent(struct mpsort_comm * c, struct mpsk_keyview kv, size_t n, uint32_t nw,
        uint64_t * d_prefix, const uint64_t * d_target, int ns, int level0, int nlevels);
/* after the stream was synchronised: aborts the job if a peer never answered */
void mps_comm_peer_descent_check(struct mpsort_comm * c);

/* ---- layout solver (mpsort_layout.c; pure host arithmetic, unit-testable) ---- */
/* C[p+1] desired cumulative output counts; clt/cle[j*(p-1) + b] local counts of rank j
 * for splitter b; nmemb[j]; writes cut[j*(p+1) + k]. Returns 0, or a negative code on
 * the reference's "serious bug" conditions (mpsort-mpi.c:707-716). */
int mpsort_solve_layout(int p, const int64_t * C, const int64_t * clt, const int64_t * cle,
        const int64_t * nmemb, int64_t * cut);
int mpsort_solve_layout2(int psrc, int pdst, const int64_t * C, const int64_t * clt, const int64_t * cle,
        const int64_t * nmemb, int64_t * cut);

#endif
