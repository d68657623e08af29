/* Single-rank MPI stand-in so that the reference's own tree / treewalk / SPH C
 * files compile and run unmodified as one "rank" (NTask = 1).  TEST
 * INFRASTRUCTURE ONLY.  With one rank every collective is a local copy and the
 * point-to-point calls are never reached (treewalk.c:580 loops i = 1..NTask-1). */
#ifndef STUB_MPI_H
#define STUB_MPI_H
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <stddef.h>
#include <omp.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;     /* value = element size in bytes */
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
typedef ptrdiff_t MPI_Aint;

#define MPI_COMM_WORLD 1
#define MPI_SUCCESS 0
#define MPI_BYTE 1
#define MPI_CHAR 1
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_UINT 4
#define MPI_FLOAT 4
#define MPI_LONG 8
#define MPI_UNSIGNED_LONG 8
#define MPI_LONG_LONG 8
#define MPI_DOUBLE 8
#define MPI_INT64_T 8
#define MPI_UINT64_T 8
#define MPI_DATATYPE_NULL 0
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_LOR 4
#define MPI_IN_PLACE ((void *) 1)
#define MPI_STATUS_IGNORE ((MPI_Status *) 0)
#define MPI_STATUSES_IGNORE ((MPI_Status *) 0)
#define MPI_REQUEST_NULL 0
#define MPI_UNDEFINED (-32766)
#define MPI_THREAD_FUNNELED 1
#define MPI_MAX_PROCESSOR_NAME 64

static inline int stub_mpi_copy(const void *s, void *r, int count, MPI_Datatype t)
{ if(s != MPI_IN_PLACE && s != r) memcpy(r, s, (size_t) count * (size_t) t); return 0; }

static inline int MPI_Init(int *a, char ***b) { return 0; }
static inline int MPI_Init_thread(int *a, char ***b, int req, int *prov) { if(prov) *prov = req; return 0; }
static inline int MPI_Finalize(void) { return 0; }
extern int ref_stub_thistask;      /* 0 unless a fixture plays another rank, see ref_stub_ntask below */
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { *r = ref_stub_thistask; return 0; }
/* 1 everywhere, except while a fixture drives a pure-computation routine of the reference for several tasks
 * (oracle/ref_domain_driver.c sets ref_stub_ntask around the call) */
extern int ref_stub_ntask;
static inline int MPI_Comm_size(MPI_Comm c, int *s) { *s = ref_stub_ntask; return 0; }
static inline int MPI_Barrier(MPI_Comm c) { return 0; }
static inline int MPI_Abort(MPI_Comm c, int e) { fprintf(stderr, "MPI_Abort(%d)\n", e); abort(); return 0; }
static inline double MPI_Wtime(void) { return omp_get_wtime(); }
static inline int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) { return 0; }
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, int root, MPI_Comm c) { return stub_mpi_copy(s, r, n, t); }
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op o, MPI_Comm c) { return stub_mpi_copy(s, r, n, t); }
static inline int MPI_Alltoall(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c) { return stub_mpi_copy(s, r, sn, st); }
static inline int MPI_Allgather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, MPI_Comm c) { return stub_mpi_copy(s, r, sn, st); }
static inline int MPI_Gather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c) { return stub_mpi_copy(s, r, sn, st); }
static inline int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype st, void *r, const int *rc, const int *rd, MPI_Datatype rt, MPI_Comm c)
{ if(s != MPI_IN_PLACE) memcpy((char *) r + (size_t) rd[0] * rt, (const char *) s + (size_t) sd[0] * st, (size_t) sc[0] * st); return 0; }
static inline int MPI_Allgatherv(const void *s, int sn, MPI_Datatype st, void *r, const int *rc, const int *rd, MPI_Datatype rt, MPI_Comm c)
{ if(s != MPI_IN_PLACE) memcpy((char *) r + (size_t) rd[0] * rt, s, (size_t) sn * st); return 0; }
static inline int MPI_Type_contiguous(int n, MPI_Datatype old, MPI_Datatype *nw) { *nw = n * old; return 0; }
static inline int MPI_Type_commit(MPI_Datatype *t) { return 0; }
static inline int MPI_Type_free(MPI_Datatype *t) { return 0; }
static inline int MPI_Type_get_extent(MPI_Datatype t, MPI_Aint *lb, MPI_Aint *ext) { *lb = 0; *ext = t; return 0; }
static inline int MPI_Isend(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *r) { fprintf(stderr, "stub MPI_Isend reached\n"); abort(); return 0; }
static inline int MPI_Irecv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request *r) { fprintf(stderr, "stub MPI_Irecv reached\n"); abort(); return 0; }
static inline int MPI_Send(const void *b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) { fprintf(stderr, "stub MPI_Send reached\n"); abort(); return 0; }
static inline int MPI_Recv(void *b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status *st) { fprintf(stderr, "stub MPI_Recv reached\n"); abort(); return 0; }
static inline int MPI_Sendrecv(const void *sb, int sn, MPI_Datatype st, int d, int stag, void *rb, int rn, MPI_Datatype rt, int s, int rtag,
                               MPI_Comm c, MPI_Status *status) { fprintf(stderr, "stub MPI_Sendrecv reached\n"); abort(); return 0; }
static inline int MPI_Waitall(int n, MPI_Request *r, MPI_Status *s) { return 0; }
static inline int MPI_Waitsome(int n, MPI_Request *r, int *outcount, int *idx, MPI_Status *s) { *outcount = MPI_UNDEFINED; return 0; }
static inline int MPI_Wait(MPI_Request *r, MPI_Status *s) { return 0; }
static inline int MPI_Test(MPI_Request *r, int *flag, MPI_Status *s) { *flag = 1; return 0; }
static inline int MPI_Igather(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c, MPI_Request *rq) { *rq = 0; return stub_mpi_copy(s, r, sn, st); }
static inline int MPI_Scatter(const void *s, int sn, MPI_Datatype st, void *r, int rn, MPI_Datatype rt, int root, MPI_Comm c) { return stub_mpi_copy(s, r, sn, st); }
static inline int MPI_Comm_free(MPI_Comm *c) { return 0; }
static inline int MPI_Cart_get(MPI_Comm c, int nd, int *dims, int *per, int *coords) { for(int i = 0; i < nd; i++) { dims[i] = 1; per[i] = 1; coords[i] = 0; } return 0; }
static inline int MPI_Cart_rank(MPI_Comm c, const int *coords, int *rank) { *rank = 0; return 0; }
static inline int MPI_Get_processor_name(char *n, int *l) { strcpy(n, "localhost"); *l = 9; return 0; }
#endif
