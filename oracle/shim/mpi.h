/* oracle/shim/mpi.h -- TEST INFRASTRUCTURE ONLY.
 * A single-process stand-in for the dozen MPI entry points FoamYade.C calls.
 * The harness (oracle/ref_harness.cpp) implements them: it plays the Yade side
 * of the wire protocol (world ranks 0..Y-1) and logs every call so the message
 * sequence of SURVEY.md section 4 can be asserted. */
#ifndef ORACLE_SHIM_MPI_H
#define ORACLE_SHIM_MPI_H
#ifdef __cplusplus
extern "C" {
#endif
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_COMM_FOAM_SHIM 1
#define MPI_DOUBLE 8
#define MPI_INT 4
#define MPI_MAX 1
#define MPI_SUM 2
int MPI_Comm_rank(MPI_Comm, int*);
int MPI_Comm_size(MPI_Comm, int*);
int MPI_Isend(const void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*);
int MPI_Wait(MPI_Request*, MPI_Status*);
int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*);
int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm);
int MPI_Allreduce(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Finalize(void);
#ifdef __cplusplus
}
#endif
#endif
