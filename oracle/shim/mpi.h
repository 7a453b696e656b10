/* oracle/shim/mpi.h -- TEST INFRASTRUCTURE ONLY.
 * One-rank MPI stand-in so the unmodified reference sources compile and run serially
 * (the reference's own tests run ./solver without mpirun: tests/LPsolver_tests:22).
 * Surface = exactly what `grep MPI_ source/` finds. */
#ifndef LP_ORACLE_SHIM_MPI_H
#define LP_ORACLE_SHIM_MPI_H
#include <stdio.h>
#include <stdlib.h>
#include <omp.h>
typedef struct { int src, tag, err; } MPI_Status;
typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3
static inline int MPI_Init_thread(int *, char ***, int required, int *provided) { *provided = required; return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
static inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm)
{ fprintf(stderr, "mpi shim: MPI_Send reached with one rank\n"); abort(); return 1; }
static inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *)
{ fprintf(stderr, "mpi shim: MPI_Recv reached with one rank\n"); abort(); return 1; }
static inline double MPI_Wtime(void) { return omp_get_wtime(); }
#endif
