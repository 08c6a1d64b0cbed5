/*
 * mpi.h -- an MPI FACADE over mpsort_comm_t (TEST / INTEGRATION GLUE, not product code).
 *
 * Purpose: the reference's own C drivers -- main-mpi.c and bench-mpi.c, compiled UNMODIFIED where
 * they lie under /root/reference, including the reference's own mpsort.h and mp-mpiu.h next to
 * them -- link against libmpsort-b200.so. With
 *
 *     typedef struct mpsort_comm * MPI_Comm;
 *
 * the reference's declarations (mpsort.h:25-47: mpsort_mpi_impl / mpsort_mpi_newarray_impl taking an
 * MPI_Comm; mp-mpiu.h:10: MPIU_Set_verbose_malloc(MPI_Comm)) are, type for type, the symbols the
 * product library exports (include/mpsort.h), so the drop-in claim is checked by the C compiler and
 * the linker. The handful of MPI calls the DRIVERS make for themselves (rank/size, a sum, a
 * neighbour send/recv, a barrier, a clock) are implemented in mpi_facade.c over the product's
 * small host collectives. No MPI exists in this image or on the GPU box.
 *
 * Ranks are either processes (RANK / WORLD_SIZE / LOCAL_RANK in the environment, one GPU each,
 * NCCL: tests/dropin/launch.py) or threads of one process on one GPU
 * (MPSORT_DROPIN_THREADS=P: the driver's main() runs once per rank thread).
 */
#ifndef MPSORT_DROPIN_MPI_H
#define MPSORT_DROPIN_MPI_H

#include <stddef.h>

#define MPI_VERSION 3
#define MPI_SUBVERSION 1
#define MPI_SUCCESS 0

struct mpsort_comm;
typedef struct mpsort_comm * MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct MPI_Status { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; } MPI_Status;

MPI_Comm mpsort_dropin_world(void);
#define MPI_COMM_WORLD (mpsort_dropin_world())

#define MPI_BYTE 1
#define MPI_INT 3
#define MPI_LONG 4
#define MPI_LONG_LONG 5
#define MPI_DOUBLE 7
#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3
#define MPI_IN_PLACE ((void *) -1)
#define MPI_STATUS_IGNORE ((MPI_Status *) 0)

int MPI_Init(int * argc, char *** argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Comm_rank(MPI_Comm comm, int * rank);
int MPI_Comm_size(MPI_Comm comm, int * size);
int MPI_Barrier(MPI_Comm comm);
double MPI_Wtime(void);
int MPI_Allreduce(const void * send, void * recv, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Bcast(void * buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Send(const void * buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void * buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm, MPI_Status * st);
int MPI_Sendrecv(const void * sbuf, int scount, MPI_Datatype stype, int dest, int stag,
                 void * rbuf, int rcount, MPI_Datatype rtype, int src, int rtag, MPI_Comm comm, MPI_Status * st);

/* thread mode: the driver is compiled with -Dmain=mpsort_dropin_main and the facade owns main() */
int mpsort_dropin_main(int argc, char ** argv);

#endif
