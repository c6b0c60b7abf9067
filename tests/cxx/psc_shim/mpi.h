// TEST INFRASTRUCTURE: a declaration-only stand-in for <mpi.h> (not installed in this image) so that the
// reference's own particles.hxx / particles_simple.hxx can be COMPILED (never linked or run) against
// include/psc_b200/psc_adapters_b200.hxx.  Nothing here is called.
#pragma once
typedef int MPI_Comm; typedef int MPI_Datatype; typedef int MPI_Op; typedef int MPI_Request; typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
typedef int MPI_Info; typedef long MPI_Aint; typedef int MPI_Group; typedef int MPI_Errhandler; typedef int MPI_File; typedef long long MPI_Offset;
#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL -1
#define MPI_COMM_SELF 1
#define MPI_SUCCESS 0
#ifdef __cplusplus
extern "C" {
#endif
int MPI_Comm_rank(MPI_Comm, int*); int MPI_Comm_size(MPI_Comm, int*); int MPI_Barrier(MPI_Comm); int MPI_Abort(MPI_Comm, int);
double MPI_Wtime(void);
int MPI_Allreduce(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Reduce(const void*, void*, int, MPI_Datatype, MPI_Op, int, MPI_Comm);
int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Gather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Gatherv(const void*, int, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, int, MPI_Comm);
int MPI_Allgather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm);
int MPI_Isend(const void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*);
int MPI_Irecv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*);
int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*);
int MPI_Waitall(int, MPI_Request*, MPI_Status*); int MPI_Wait(MPI_Request*, MPI_Status*);
int MPI_Comm_dup(MPI_Comm, MPI_Comm*); int MPI_Comm_free(MPI_Comm*);
#ifdef __cplusplus
}
#endif
#define MPI_INT 1
#define MPI_DOUBLE 2
#define MPI_FLOAT 3
#define MPI_UNSIGNED 4
#define MPI_LONG 5
#define MPI_BYTE 6
#define MPI_CHAR 7
#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_SUM 3
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_REQUEST_NULL 0
#define MPI_IN_PLACE ((void*)1)
