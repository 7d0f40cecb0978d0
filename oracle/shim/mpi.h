/* TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
 *
 * Single-rank stand-in for the handful of MPI entry points the reference's
 * cedr/ sources reference, so they compile without an MPI installation. With
 * one rank every collective is a copy, and point-to-point calls are never
 * reached (a single-rank tree::analyze produces no comm partners).
 */
#ifndef CEDR_B200_ORACLE_MPI_SHIM_H
#define CEDR_B200_ORACLE_MPI_SHIM_H

#include <cstring>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
struct MPI_Status { int unused; };
typedef void MPI_User_function(void*, void*, int*, MPI_Datatype*);

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_STATUS_IGNORE (static_cast<MPI_Status*>(0))

/* Datatype tags encode the element size in bytes in the low byte. */
enum { MPI_INT = 0x104, MPI_DOUBLE = 0x208, MPI_LONG_INT = 0x308 };
enum { MPI_SUM = 1, MPI_MAX, MPI_MIN, MPI_LAND };

static inline size_t mpi_shim_bytes (int count, MPI_Datatype dt) {
  return static_cast<size_t>(count)*static_cast<size_t>(dt & 0xff);
}

static inline int MPI_Init (int*, char***) { return MPI_SUCCESS; }
static inline int MPI_Finalize () { return MPI_SUCCESS; }
static inline int MPI_Finalized (int* flag) { *flag = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_rank (MPI_Comm, int* rank) { *rank = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size (MPI_Comm, int* size) { *size = 1; return MPI_SUCCESS; }
static inline int MPI_Barrier (MPI_Comm) { return MPI_SUCCESS; }

static inline int MPI_Allreduce (const void* s, void* r, int n, MPI_Datatype dt,
                                 MPI_Op, MPI_Comm) {
  std::memmove(r, s, mpi_shim_bytes(n, dt));
  return MPI_SUCCESS;
}
static inline int MPI_Reduce (const void* s, void* r, int n, MPI_Datatype dt,
                              MPI_Op, int, MPI_Comm) {
  std::memmove(r, s, mpi_shim_bytes(n, dt));
  return MPI_SUCCESS;
}
static inline int MPI_Gather (const void* s, int n, MPI_Datatype dt, void* r, int,
                              MPI_Datatype, int, MPI_Comm) {
  std::memmove(r, s, mpi_shim_bytes(n, dt));
  return MPI_SUCCESS;
}
static inline int MPI_Gatherv (const void* s, int n, MPI_Datatype dt, void* r,
                               const int*, const int*, MPI_Datatype, int, MPI_Comm) {
  std::memmove(r, s, mpi_shim_bytes(n, dt));
  return MPI_SUCCESS;
}

/* Never reached with one rank; return non-success so a surprise is loud. */
static inline int MPI_Isend (const void*, int, MPI_Datatype, int, int, MPI_Comm,
                             MPI_Request*) { return 1; }
static inline int MPI_Irecv (void*, int, MPI_Datatype, int, int, MPI_Comm,
                             MPI_Request*) { return 1; }
static inline int MPI_Request_free (MPI_Request*) { return MPI_SUCCESS; }
static inline int MPI_Waitall (int, MPI_Request*, MPI_Status*) { return MPI_SUCCESS; }
static inline int MPI_Waitany (int, MPI_Request*, int* idx, MPI_Status*) {
  *idx = 0;
  return MPI_SUCCESS;
}
static inline int MPI_Op_create (MPI_User_function*, int, MPI_Op* op) {
  *op = 99;
  return MPI_SUCCESS;
}
static inline int MPI_Op_free (MPI_Op*) { return MPI_SUCCESS; }

#endif
