/* libxsmm_b200 -- fixed-size sparse (A) x dense (B) multiplication, row-major:
 *   C(M x N, ldc) = beta * C + alpha * A(M x K, lda; given dense, zeros dropped) * B(K x N, ldb)
 * API and semantics follow the reference include/libxsmm_fsspmdm.h:26-45 and
 * src/libxsmm_fsspmdm.c:24-560: beta in {0,1}, F32/F64, N % (64/sizeof(T)) == 0, lda >= K,
 * ldb >= N, ldc >= N, NULL for an all-zero A. The sparsity pattern and alpha-scaled values are
 * frozen at create time and kept in device memory; execute() launches one streaming kernel.
 */
#ifndef LIBXSMM_FSSPMDM_H
#define LIBXSMM_FSSPMDM_H

#include "libxsmm_typedefs.h"

#define libxsmm_dfsspmdm libxsmm_fsspmdm
#define libxsmm_sfsspmdm libxsmm_fsspmdm
typedef struct libxsmm_fsspmdm libxsmm_fsspmdm;

#if defined(__cplusplus)
extern "C" {
#endif
LIBXSMM_API libxsmm_fsspmdm* libxsmm_fsspmdm_create(libxsmm_datatype datatype,
  libxsmm_blasint M, libxsmm_blasint N, libxsmm_blasint K, libxsmm_blasint lda, libxsmm_blasint ldb, libxsmm_blasint ldc,
  const void* alpha, const void* beta, const void* a_dense, int LIBXSMM_ARGDEF(c_is_nt, 0),
  libxsmm_timer_tickint LIBXSMM_ARGDEF((*timer_tick)(void), NULL));
LIBXSMM_API libxsmm_dfsspmdm* libxsmm_dfsspmdm_create(
  libxsmm_blasint M, libxsmm_blasint N, libxsmm_blasint K, libxsmm_blasint lda, libxsmm_blasint ldb, libxsmm_blasint ldc,
  double alpha, double beta, const double* a_dense, int LIBXSMM_ARGDEF(c_is_nt, 0),
  libxsmm_timer_tickint LIBXSMM_ARGDEF((*timer_tick)(void), NULL));
LIBXSMM_API libxsmm_sfsspmdm* libxsmm_sfsspmdm_create(
  libxsmm_blasint M, libxsmm_blasint N, libxsmm_blasint K, libxsmm_blasint lda, libxsmm_blasint ldb, libxsmm_blasint ldc,
  float alpha, float beta, const float* a_dense, int LIBXSMM_ARGDEF(c_is_nt, 0),
  libxsmm_timer_tickint LIBXSMM_ARGDEF((*timer_tick)(void), NULL));

LIBXSMM_API void libxsmm_fsspmdm_execute(const libxsmm_fsspmdm* handle, const void* B, void* C);
LIBXSMM_API void libxsmm_dfsspmdm_execute(const libxsmm_dfsspmdm* handle, const double* B, double* C);
LIBXSMM_API void libxsmm_sfsspmdm_execute(const libxsmm_sfsspmdm* handle, const float* B, float* C);

LIBXSMM_API void libxsmm_fsspmdm_destroy(libxsmm_fsspmdm* handle);
LIBXSMM_API void libxsmm_dfsspmdm_destroy(libxsmm_dfsspmdm* handle);
LIBXSMM_API void libxsmm_sfsspmdm_destroy(libxsmm_sfsspmdm* handle);
#if defined(__cplusplus)
}
#endif
#endif /* LIBXSMM_FSSPMDM_H */
