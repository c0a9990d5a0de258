/* Single-rank stand-in for <mpi.h>.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference's lib/ includes "mpi.h" unconditionally (lib/grid.h:10,
 * lib/edm_bias.cpp:12) and names MPI symbols outside its EDM_SERIAL guards
 * (lib/grid.h:523-647, lib/edm_bias.cpp:620-905).  No MPI exists in this image, so the
 * oracle build (oracle/Makefile) puts this directory on the include path.  Every call
 * behaves as a communicator of exactly one rank: queries return rank 0 / size 1,
 * reductions and gathers copy send -> recv, everything else is a no-op.
 */
#ifndef EDM_ORACLE_STUB_MPI_H
#define EDM_ORACLE_STUB_MPI_H

#include <string.h>
#include <stddef.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int unused; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_INT 4
#define MPI_UNSIGNED 5
#define MPI_DOUBLE 8
#define MPI_MAX 1
#define MPI_SUM 2
#define MPI_STATUS_IGNORE ((MPI_Status*)0)

static inline size_t edm_stub_mpi_width(MPI_Datatype t) {
  return t == MPI_DOUBLE ? sizeof(double) : sizeof(int);
}
static inline int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int* s) { (void)c; *s = 1; return 0; }
static inline int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c) {
  (void)o; (void)c; memcpy(r, s, (size_t)n * edm_stub_mpi_width(t)); return 0;
}
static inline int MPI_Gather(const void* s, int n, MPI_Datatype t, void* r, int rn, MPI_Datatype rt,
                             int root, MPI_Comm c) {
  (void)rn; (void)rt; (void)root; (void)c; memcpy(r, s, (size_t)n * edm_stub_mpi_width(t)); return 0;
}
static inline int MPI_Scatter(const void* s, int n, MPI_Datatype t, void* r, int rn, MPI_Datatype rt,
                              int root, MPI_Comm c) {
  (void)rn; (void)rt; (void)root; (void)c; memcpy(r, s, (size_t)n * edm_stub_mpi_width(t)); return 0;
}
static inline int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) {
  (void)b; (void)n; (void)t; (void)root; (void)c; return 0;
}
static inline int MPI_Isend(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request* q) {
  (void)b; (void)n; (void)t; (void)d; (void)tag; (void)c; *q = 0; return 0;
}
static inline int MPI_Recv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status* st) {
  (void)b; (void)n; (void)t; (void)s; (void)tag; (void)c; (void)st; return 0;
}
static inline int MPI_Wait(MPI_Request* q, MPI_Status* st) { (void)q; (void)st; return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }

#endif
