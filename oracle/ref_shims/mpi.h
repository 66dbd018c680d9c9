/* Single-rank stand-in for <mpi.h>, used only to compile the reference's GMG sources
 * as a CPU oracle (oracle/_ref).  TEST INFRASTRUCTURE - never linked into the product.
 * Semantics on one rank: reductions/scans are copies, rank 0 of 1, p2p is never reached. */
#ifndef ORACLE_SHIM_MPI_H
#define ORACLE_SHIM_MPI_H
#include <cstddef>
#include <cstring>
#include <cstdlib>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR, count; };
#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_INT 4
#define MPI_DOUBLE 8
#define MPI_BYTE 1
#define MPI_CHAR 1
#define MPI_SUM 0
#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_STATUSES_IGNORE ((MPI_Status *) 0)
#define MPI_STATUS_IGNORE ((MPI_Status *) 0)
#define MPI_SUCCESS 0
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline int MPI_Allreduce(const void *in, void *out, int n, MPI_Datatype t, MPI_Op, MPI_Comm)
{ std::memcpy(out, in, (size_t) n * (size_t) t); return 0; }
static inline int MPI_Scan(const void *in, void *out, int n, MPI_Datatype t, MPI_Op, MPI_Comm)
{ std::memcpy(out, in, (size_t) n * (size_t) t); return 0; }
static inline int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { std::abort(); }
static inline int MPI_Probe(int, int, MPI_Comm, MPI_Status *) { std::abort(); }
static inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { std::abort(); }
static inline int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *) { std::abort(); }
static inline int MPI_Waitall(int n, MPI_Request *, MPI_Status *) { if (n) std::abort(); return 0; }
static inline int MPI_Init(int *, char ***) { return 0; }
static inline int MPI_Finalize() { return 0; }
#endif
