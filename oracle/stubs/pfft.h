/* Single-rank stand-in for PFFT 1.0.8-alpha3 (absent offline; fetched by the reference's
 * depends/install_pfft.sh at build time).  TEST INFRASTRUCTURE ONLY.
 *
 * Declares exactly what libgadget/petapm.c uses (petapm.c:95-97,137,147-187,228-229,305,344);
 * the functions are implemented in oracle/pfft_standin.c as plain separable DFTs with PFFT's
 * conventions for one process: unnormalised transforms, FFTW signs (forward e^{-i}), real input
 * [x][y][z] contiguous, PFFT_TRANSPOSED_OUT half spectrum stored [y][z][x] with z in [0, N/2]. */
#ifndef STUB_PFFT_H
#define STUB_PFFT_H
#include <stddef.h>
#include <mpi.h>
typedef double pfft_complex[2];
typedef struct pfft_plan_s *pfft_plan;
#define PFFT_FORWARD (-1)
#define PFFT_BACKWARD (+1)
#define PFFT_TRANSPOSED_NONE 0u
#define PFFT_TRANSPOSED_IN 1u
#define PFFT_TRANSPOSED_OUT 2u
#define PFFT_ESTIMATE 4u
#define PFFT_TUNE 8u
#define PFFT_DESTROY_INPUT 16u
void pfft_init(void);
void pfft_plan_with_nthreads(int nthreads);
int pfft_create_procmesh_2d(MPI_Comm comm, int np0, int np1, MPI_Comm *comm_cart_2d);
ptrdiff_t pfft_local_size_dft_r2c_3d(const ptrdiff_t *n, MPI_Comm comm_cart, unsigned pfft_flags,
                                     ptrdiff_t *local_ni, ptrdiff_t *local_i_start, ptrdiff_t *local_no, ptrdiff_t *local_o_start);
pfft_plan pfft_plan_dft_r2c_3d(const ptrdiff_t *n, double *in, pfft_complex *out, MPI_Comm comm_cart, int sign, unsigned pfft_flags);
pfft_plan pfft_plan_dft_c2r_3d(const ptrdiff_t *n, pfft_complex *in, double *out, MPI_Comm comm_cart, int sign, unsigned pfft_flags);
void pfft_execute_dft_r2c(const pfft_plan plan, double *in, pfft_complex *out);
void pfft_execute_dft_c2r(const pfft_plan plan, pfft_complex *in, double *out);
void pfft_destroy_plan(pfft_plan plan);
#endif
