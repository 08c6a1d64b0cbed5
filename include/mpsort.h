/*
 * mpsort.h -- C ABI of mpsort-b200: a B200-native (sm_100a CUDA + NCCL) distributed
 * histogram sort that is a drop-in for MP-sort's `mpsort_mpi` hot path.
 *
 * Every entry point below names the reference interface it replaces
 * (paths relative to the MP-sort v0.1.19 source tree).
 *
 * Differences a caller sees, and why:
 *   - the host `radix()` callback + `rsize` + `arg` triple (reference mpsort.h:25-43)
 *     becomes a device-side key descriptor, `struct mpsort_radix_desc`: a host
 *     function pointer cannot run on the GPU.
 *   - `MPI_Comm` becomes the opaque `mpsort_comm_t` (rank, size, NCCL communicator,
 *     CUDA stream, device arena): there is no MPI on the target.
 *   - `base` / `out` may be host pointers (pageable or pinned) or device pointers;
 *     the library detects which with cudaPointerGetAttributes.
 * Everything else -- names, option bits, env variables, in-place semantics,
 * error convention (message with the caller's file:line on stderr, then abort),
 * collectivity, timers/report -- follows the reference.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point aborts.
 */
#ifndef MPSORT_B200_H
#define MPSORT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPSORT_B200_VERSION "0.1.0"

/* ------------------------------------------------------------------------- */
/* Option bits: identical values to reference mpsort.h:17-20.                */
#define MPSORT_DISABLE_SPARSE_ALLTOALLV (1 << 1)
#define MPSORT_DISABLE_GATHER_SORT      (1 << 3)
#define MPSORT_REQUIRE_GATHER_SORT      (1 << 4)
#define MPSORT_REQUIRE_SPARSE_ALLTOALLV (1 << 6)
/* Extension bit (not in the reference): recompute the reference's byte checksum
 * (mpsort-mpi.c:148-159,193,324-330) on the device before and after the sort and
 * abort on mismatch. Also enabled by env MPSORT_VERIFY_CHECKSUM. Off by default:
 * it is two extra full passes over the data and never changes the output. */
#define MPSORT_VERIFY_CHECKSUM          (1 << 8)

/* reference mpsort.h:22-24 / mpsort-mpi.c:747-767. Process-global, lazily OR'd with
 * env MPSORT_DISABLE_SPARSE_ALLTOALLV, MPSORT_DISABLE_GATHER_SORT,
 * MPSORT_REQUIRE_GATHER_SORT, MPSORT_REQUIRE_SPARSE_ALLTOALLV on first use. */
void mpsort_mpi_set_options(int options);
int  mpsort_mpi_has_options(int options);
void mpsort_mpi_unset_options(int options);

/* ------------------------------------------------------------------------- */
/* Radix key descriptor: device-side equivalent of the reference's
 *   void radix(const void *ptr, void *radix, void *arg)  +  rsize
 * (mpsort.h:1-4; binding.pyx:44-121 for the Python-generated ones).
 *
 * The key of a record is `nwords` little-endian integers of `width` bytes each,
 * stored consecutively at byte `offset` inside the record. Word nwords-1 is the
 * MOST significant (radixsort.c:82-98,178-185; binding.pyx:136-137). If
 * `is_signed` every word is a two's complement integer and is biased by flipping
 * its sign bit (binding.pyx:91-100,112-121; bench-mpi.c:13-15), which makes the
 * unsigned comparison below order it correctly.
 *
 * Ordering is the reference's: unsigned comparison of the radix from its most
 * significant word down (radixsort.c:47-63 for rsize 2/4/8, :65-98 otherwise).
 * width must be 1, 2, 4 or 8; offset + width*nwords must be <= elsize.
 *
 * Mapping of reference call sites:
 *   bench-mpi.c:13-15  radix_int (int64 + INT64_MIN)  -> {0, 8, 1, 1}
 *   main-mpi.c:10-12   radix_int (raw int, rsize 4)   -> {0, 4, 1, 0}
 *   binding.pyx u8/i8/u4/i4 field f with k elements   -> {offsetof f, 8|4, k, 0|1}
 *   test-issue7.c:12-17 (8 bytes @16, 4 bytes @0)     -> not one contiguous field;
 *        use the Python-test layout (test_mpsort.py:538-541) {0, 8, 2, 0}.
 */
struct mpsort_radix_desc {
    size_t   offset;    /* byte offset of the key inside a record          */
    uint32_t width;     /* bytes per key word: 1, 2, 4 or 8                */
    uint32_t nwords;    /* number of words; the last one is most significant */
    int32_t  is_signed; /* non-zero: words are signed integers              */
    int32_t  reserved;  /* must be 0                                        */
};

/* ------------------------------------------------------------------------- */
/* Communicator: replaces MPI_Comm in every signature (mpsort.h:25-49).      */
typedef struct mpsort_comm * mpsort_comm_t;

#define MPSORT_UNIQUE_ID_BYTES 128

/* Fill `id` (MPSORT_UNIQUE_ID_BYTES bytes) on ONE rank; ship the bytes to all other
 * ranks by any out-of-band means (file, socket, torchrun's store, MPI_Bcast ...).
 * Wraps ncclGetUniqueId. Returns 0 on success. */
int mpsort_comm_get_unique_id(void * id);

/* One process per GPU (the production transport): collective over all `size`
 * ranks. Binds the calling process to CUDA device `device`, creates a stream and an
 * NCCL communicator. size == 1 never touches NCCL. Replaces MPI_Init + the MPI_Comm
 * argument. Aborts on CUDA/NCCL failure. */
mpsort_comm_t mpsort_comm_init_rank(int rank, int size, const void * unique_id, int device);

/* A size-1 communicator on `device` (the single-GPU path; no NCCL). */
mpsort_comm_t mpsort_comm_self(int device);

/* In-process rank group: `size` communicators whose ranks are driven by `size`
 * host threads of THIS process (one thread per rank, each calling the collective
 * entry points concurrently). devices[i] is the CUDA device of rank i; several
 * ranks may share one device. Records move by peer/device-to-device copies instead
 * of NCCL (NCCL refuses two ranks on one device). This is how the reference's 4- and
 * 12-rank tests run on a box with fewer GPUs, and how a single-process multi-GPU
 * caller uses the library. Returns 0 on success and fills comms[0..size). */
int mpsort_comm_init_local_group(int size, const int * devices, mpsort_comm_t * comms);

void mpsort_comm_destroy(mpsort_comm_t comm);
int  mpsort_comm_rank(mpsort_comm_t comm);
int  mpsort_comm_size(mpsort_comm_t comm);
int  mpsort_comm_device(mpsort_comm_t comm);
/* The CUDA stream (cudaStream_t) all device work of this communicator runs on. */
void * mpsort_comm_stream(mpsort_comm_t comm);

/* Small host-side collectives over the communicator, for callers that have no MPI:
 * the Python layer builds comm.allgather/allreduce/bcast/barrier on these
 * (replaces the mpi4py calls in binding.pyx:182-183 and test_mpsort.py:8-20). */
void mpsort_comm_barrier(mpsort_comm_t comm);
void mpsort_comm_allgather_host(mpsort_comm_t comm, const void * send, void * recv, size_t nbytes_per_rank);
/* Variable-size version: recvcounts[size] bytes from each rank, packed in rank order. */
void mpsort_comm_allgatherv_host(mpsort_comm_t comm, const void * send, size_t nbytes,
                                 void * recv, const size_t * recvcounts);

/* ------------------------------------------------------------------------- */
/* The sort. Replaces mpsort_mpi_impl (mpsort.h:25-29, mpsort-mpi.c:129-142).
 * Collective. In place: on return rank k's `base` holds the k-th chunk (of its
 * original length nmemb) of the global stable sort of all ranks' records. */
void mpsort_mpi_desc_impl(void * base, size_t nmemb, size_t elsize,
        const struct mpsort_radix_desc * desc,
        mpsort_comm_t comm,
        const int line, const char * file);

#define mpsort_mpi_desc(base, nmemb, elsize, desc, comm) \
    mpsort_mpi_desc_impl(base, nmemb, elsize, desc, comm, __LINE__, __FILE__)

/* Replaces mpsort_mpi_newarray_impl (mpsort.h:35-42, mpsort-mpi.c:161-331).
 * `out` may alias `base` exactly (in place) or be a distinct buffer of `outnmemb`
 * records; sum(outnmemb) over ranks must equal sum(nmemb) or the job aborts
 * (mpsort-mpi.c:225-232). With out != base the contents of `base` after the call
 * are unspecified (the reference leaves it locally sorted; not contractual). */
void mpsort_mpi_newarray_desc_impl(void * base, size_t nmemb,
        void * out, size_t outnmemb,
        size_t elsize,
        const struct mpsort_radix_desc * desc,
        mpsort_comm_t comm,
        const int line, const char * file);

#define mpsort_mpi_newarray_desc(base, nmemb, out, outnmemb, elsize, desc, comm) \
    mpsort_mpi_newarray_desc_impl(base, nmemb, out, outnmemb, elsize, desc, comm, \
    __LINE__, __FILE__)

/* Replaces radix_sort (mpsort.h:1-4, radixsort.c:35-44): stable local sort of one
 * array on one GPU. `device` is the CUDA device ordinal. */
void radix_sort_desc(void * base, size_t nmemb, size_t size,
        const struct mpsort_radix_desc * desc, int device);

/* radix_sort_desc and radix_sort have no communicator argument: they use a size-1 communicator
 * (stream + grow-only device arena, several times nmemb * size after a large call) cached per
 * calling thread and device. mpsort_release_cached() destroys the calling thread's cached
 * communicators and gives their device memory back; a thread that exits releases its own. */
void mpsort_release_cached(void);

/* ------------------------------------------------------------------------- */
/* The reference's OWN signatures (mpsort.h:1-4, :25-47), for callers that keep a host
 * radix() callback: e.g. test-issue7.c:12-17 builds its key from two separate fields,
 * which no single descriptor expresses. A host function pointer cannot run on the
 * GPU, so the callback is evaluated once per record on the host into the front of an
 * augmented record {radix | record}; the augmented records are sorted on the GPU by
 * the descriptor mpsort_callback_desc(rsize) and the records copied back. Same
 * ordering as the reference (radixsort.c:47-98,178-193: the radix compares as a
 * little-endian integer of rsize bytes), same in-place / newarray semantics, same
 * error convention. `base` / `out` must be HOST memory on this path (aborts on a
 * device pointer: use the descriptor entry points for device-resident data);
 * 1 <= rsize <= 128. Only the communicator type differs from the reference. */
typedef void (*mpsort_radix_func)(const void * ptr, void * radix, void * arg);

void mpsort_mpi_impl(void * base, size_t nmemb, size_t elsize,
        mpsort_radix_func radix, size_t rsize, void * arg,
        mpsort_comm_t comm, const int line, const char * file);
#define mpsort_mpi(base, nmemb, elsize, radix, rsize, arg, comm) \
    mpsort_mpi_impl(base, nmemb, elsize, radix, rsize, arg, comm, __LINE__, __FILE__)

void mpsort_mpi_newarray_impl(void * base, size_t nmemb,
        void * out, size_t outnmemb, size_t elsize,
        mpsort_radix_func radix, size_t rsize, void * arg,
        mpsort_comm_t comm, const int line, const char * file);
#define mpsort_mpi_newarray(base, nmemb, out, outnmemb, elsize, radix, rsize, arg, comm) \
    mpsort_mpi_newarray_impl(base, nmemb, out, outnmemb, elsize, radix, rsize, arg, comm, \
    __LINE__, __FILE__)

/* reference radix_sort (mpsort.h:1-4, radixsort.c:35-44): stable local sort of a host
 * array on the calling thread's current CUDA device. */
void radix_sort(void * base, size_t nmemb, size_t size,
        mpsort_radix_func radix, size_t rsize, void * arg);

/* The pieces of the callback path, exported so that they can be checked without a GPU:
 * the descriptor + padded radix size an rsize-byte radix sorts under (0, or -1 if rsize is
 * out of range); the host pass that writes {radix, zero padding to rpad, record} per
 * record into `aug` (nmemb * (rpad + elsize) bytes); and its inverse. */
int  mpsort_callback_desc(size_t rsize, struct mpsort_radix_desc * desc, size_t * rpad);
void mpsort_callback_pack(const void * base, size_t nmemb, size_t elsize,
        mpsort_radix_func radix, size_t rsize, void * arg, void * aug);
void mpsort_callback_unpack(const void * aug, size_t nmemb, size_t elsize, size_t rsize, void * out);

/* Replaces mpsort_mpi_report_last_run (mpsort.h:49, mpsort-mpi.c:110-119): prints
 * "<phase>: <seconds>" for FirstSort, PmaxPmin, bisectNNNN, findP, LayDistr,
 * LaySolve, Exchange, SecondSort, measured with CUDA events on the comm's stream. */
void mpsort_mpi_report_last_run(void);

/* Programmatic access to the same timers. Returns the number of phases and, for
 * i < max, stores a pointer to a static phase name and its duration in seconds. */
int mpsort_mpi_get_last_run(const char ** names, double * seconds, int max);

/* Statistics of the last call on this communicator (for benches and parity tests).
 * sendcounts receives this rank's row of the exchange matrix (items sent to each
 * rank; reference SendCount[] mpsort-mpi.c:483-485), up to `max` entries. */
struct mpsort_last_stats {
    uint64_t nmemb, outnmemb, elsize;
    uint32_t key_words;          /* 64-bit words the key was packed into        */
    uint32_t first_sort_passes;  /* 8-bit radix passes executed (constant digits skipped) */
    uint32_t second_sort_passes;
    uint32_t splitter_rounds;    /* device count+allreduce rounds               */
    uint32_t used_gather;        /* 1 if the small-input gather path was taken  */
    uint32_t dense_exchange;     /* 1 if zero-length pairs were also posted     */
    uint64_t bytes_sent_remote;  /* record bytes that left this GPU             */
    uint32_t second_sort_merge_tiles; /* > 0: SecondSort ran as a p-way merge of this many tiles */
    uint32_t record_mode;        /* 1: 16-byte records were carried through the passes themselves */
    uint32_t hybrid;             /* 1: four high-digit passes + run fix-up instead of all passes */
    uint32_t hybrid_long_runs;   /* runs of > 256 equal high parts that were sorted separately */
    uint32_t rebased;            /* 1: keys were sorted relative to their minimum (fewer passes) */
    uint32_t p2p_exchange;       /* 1: records moved by peer stores (CUDA IPC), 0: ncclSend/ncclRecv or copies */
    uint32_t exchange_phases;    /* parts the exchange + merge were pipelined in (1 = not pipelined) */
    uint32_t own_slices_in_place;/* parts whose own slice was merged from the send buffer, never copied */
};
void mpsort_comm_last_stats(mpsort_comm_t comm, struct mpsort_last_stats * st,
                            int64_t * sendcounts, int max);

/* ------------------------------------------------------------------------- */
/* Allocator hook for HOST allocations of significance, as in the reference
 * (mp-mpiu.h:4-19, mp-mpiu.c:9-59). Device memory comes from a per-communicator
 * grow-only arena and is reported through the verbose hook. */
typedef void * (*mpiu_malloc_func)(const char * name, size_t size, const char * file, const int line, void * userdata);
typedef void (*mpiu_free_func)(void * ptr, const char * file, const int line, void * userdata);
void mpiu_set_malloc(mpiu_malloc_func malloc_func, mpiu_free_func free_func, void * userdata);
#define MPIU_SetMalloc mpiu_set_malloc
void MPIU_Set_verbose_malloc(mpsort_comm_t comm);

#ifdef __cplusplus
}
#endif
#endif
