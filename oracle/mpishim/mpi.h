/*
 * mpi.h -- single-host MPI shim (TEST INFRASTRUCTURE, not product code).
 *
 * Neither this container nor the GPU box has an MPI. The reference's MPI sources
 * (mpsort-mpi.c, mp-mpiu.c, bench-mpi.c, main-mpi.c) compile UNMODIFIED against
 * this header and link against mpishim.c, which implements the ~30 MPI calls they
 * use over one shared-memory file, one PROCESS per rank
 * (the reference keeps per-process statics: mpsort-mpi.c:17,33,105-108).
 * Launch with `mpirun-shim -np N prog args...`; without the launcher a program
 * runs as a singleton (size 1).
 *
 * The shim contains no sort arithmetic: it only moves bytes and sums integers.
 */
#ifndef MPISHIM_MPI_H
#define MPISHIM_MPI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPI_VERSION 3
#define MPI_SUBVERSION 1
#define MPISHIM 1

#define MPI_SUCCESS 0

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef ptrdiff_t MPI_Aint;
typedef struct MPI_Status { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; } MPI_Status;

#define MPI_COMM_NULL   0
#define MPI_COMM_WORLD  1
#define MPI_COMM_SELF   2

/* datatypes: non-zero (mpsort-mpi.c:33,173 tests `== 0` for "unset") */
#define MPI_DATATYPE_NULL (-1)
#define MPI_BYTE       1
#define MPI_CHAR       2
#define MPI_INT        3
#define MPI_LONG       4
#define MPI_LONG_LONG  5
#define MPI_UNSIGNED_LONG 6
#define MPI_DOUBLE     7
/* derived contiguous types carry their extent: 0x40000000 | nbytes */
#define MPISHIM_DERIVED 0x40000000

#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3

#define MPI_IN_PLACE ((void *) -1)
#define MPI_STATUS_IGNORE ((MPI_Status *) 0)
#define MPI_STATUSES_IGNORE ((MPI_Status *) 0)
#define MPI_REQUEST_NULL (-1)
#define MPI_UNDEFINED (-32766)

int MPI_Init(int * argc, char *** argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
double MPI_Wtime(void);

int MPI_Comm_size(MPI_Comm comm, int * size);
int MPI_Comm_rank(MPI_Comm comm, int * rank);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm * newcomm);
int MPI_Comm_free(MPI_Comm * comm);

int MPI_Type_contiguous(int count, MPI_Datatype oldtype, MPI_Datatype * newtype);
int MPI_Type_commit(MPI_Datatype * type);
int MPI_Type_free(MPI_Datatype * type);
int MPI_Type_get_extent(MPI_Datatype type, MPI_Aint * lb, MPI_Aint * extent);

int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void * buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Allreduce(const void * sendbuf, void * recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Allgather(const void * sendbuf, int sendcount, MPI_Datatype sendtype,
                  void * recvbuf, int recvcount, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Alltoall(const void * sendbuf, int sendcount, MPI_Datatype sendtype,
                 void * recvbuf, int recvcount, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Alltoallv(const void * sendbuf, const int * sendcounts, const int * sdispls, MPI_Datatype sendtype,
                  void * recvbuf, const int * recvcounts, const int * rdispls, MPI_Datatype recvtype, MPI_Comm comm);
int MPI_Gather(const void * sendbuf, int sendcount, MPI_Datatype sendtype,
               void * recvbuf, int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Gatherv(const void * sendbuf, int sendcount, MPI_Datatype sendtype,
                void * recvbuf, const int * recvcounts, const int * displs, MPI_Datatype recvtype,
                int root, MPI_Comm comm);
int MPI_Scatterv(const void * sendbuf, const int * sendcounts, const int * displs, MPI_Datatype sendtype,
                 void * recvbuf, int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm);

int MPI_Send(const void * buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void * buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status * status);
int MPI_Sendrecv(const void * sendbuf, int sendcount, MPI_Datatype sendtype, int dest, int sendtag,
                 void * recvbuf, int recvcount, MPI_Datatype recvtype, int source, int recvtag,
                 MPI_Comm comm, MPI_Status * status);
int MPI_Isend(const void * buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request * req);
int MPI_Irecv(void * buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request * req);
int MPI_Waitall(int count, MPI_Request * reqs, MPI_Status * statuses);
int MPI_Wait(MPI_Request * req, MPI_Status * status);

#ifdef __cplusplus
}
#endif
#endif
