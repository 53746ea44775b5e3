/* libxsmm_b200 -- fixed sparse operator times dense matrix ("fsspmdm"), row-major:
 *
 *     C[rows x cols] = beta * C + alpha * A[rows x inner] * B[inner x cols]
 *
 * A is handed over as a DENSE array once, at create time; entries with alpha*a == 0 are dropped and
 * the remaining pattern plus alpha-scaled values are frozen in device memory, execute() launches one
 * streaming kernel over B and C. Constraints (as in the reference, src/libxsmm_fsspmdm.c:24-140):
 * beta is 0 or 1, F32 or F64 only, cols % (64/sizeof(T)) == 0, ld_a >= inner, ld_b >= cols,
 * ld_c >= cols; an all-zero operator yields NULL.
 *
 * ABI-compatible with the reference's include/libxsmm_fsspmdm.h:26-45 (same symbols, same argument
 * order and types); the two trailing create arguments are accepted and ignored: `c_is_nt` (streaming
 * store hint) and `timer_tick` (the reference benchmarks alternative x86 kernels with it).
 */
#ifndef LIBXSMM_FSSPMDM_H
#define LIBXSMM_FSSPMDM_H

#include "libxsmm_typedefs.h"

typedef struct libxsmm_fsspmdm libxsmm_fsspmdm;
/* the typed front ends share the one opaque handle type */
#define libxsmm_sfsspmdm libxsmm_fsspmdm
#define libxsmm_dfsspmdm libxsmm_fsspmdm

#if defined(__cplusplus)
extern "C" {
#endif

/* ---- type-erased: alpha/beta/dense_a point to values of `precision` (F32 or F64) ---------------- */
LIBXSMM_API libxsmm_fsspmdm* libxsmm_fsspmdm_create(libxsmm_datatype precision, libxsmm_blasint rows, libxsmm_blasint cols,
                                                    libxsmm_blasint inner, libxsmm_blasint ld_a, libxsmm_blasint ld_b, libxsmm_blasint ld_c,
                                                    const void* alpha, const void* beta, const void* dense_a,
                                                    int LIBXSMM_ARGDEF(c_is_nt, 0), libxsmm_timer_tickint LIBXSMM_ARGDEF((*timer_tick)(void), NULL));
LIBXSMM_API void libxsmm_fsspmdm_execute(const libxsmm_fsspmdm* op, const void* b, void* c);
LIBXSMM_API void libxsmm_fsspmdm_destroy(libxsmm_fsspmdm* op);

/* ---- single precision --------------------------------------------------------------------------- */
LIBXSMM_API libxsmm_sfsspmdm* libxsmm_sfsspmdm_create(libxsmm_blasint rows, libxsmm_blasint cols, libxsmm_blasint inner,
                                                      libxsmm_blasint ld_a, libxsmm_blasint ld_b, libxsmm_blasint ld_c,
                                                      float alpha, float beta, const float* dense_a,
                                                      int LIBXSMM_ARGDEF(c_is_nt, 0), libxsmm_timer_tickint LIBXSMM_ARGDEF((*timer_tick)(void), NULL));
LIBXSMM_API void libxsmm_sfsspmdm_execute(const libxsmm_sfsspmdm* op, const float* b, float* c);
LIBXSMM_API void libxsmm_sfsspmdm_destroy(libxsmm_sfsspmdm* op);

/* ---- double precision --------------------------------------------------------------------------- */
LIBXSMM_API libxsmm_dfsspmdm* libxsmm_dfsspmdm_create(libxsmm_blasint rows, libxsmm_blasint cols, libxsmm_blasint inner,
                                                      libxsmm_blasint ld_a, libxsmm_blasint ld_b, libxsmm_blasint ld_c,
                                                      double alpha, double beta, const double* dense_a,
                                                      int LIBXSMM_ARGDEF(c_is_nt, 0), libxsmm_timer_tickint LIBXSMM_ARGDEF((*timer_tick)(void), NULL));
LIBXSMM_API void libxsmm_dfsspmdm_execute(const libxsmm_dfsspmdm* op, const double* b, double* c);
LIBXSMM_API void libxsmm_dfsspmdm_destroy(libxsmm_dfsspmdm* op);

#if defined(__cplusplus)
}
#endif
#endif /* LIBXSMM_FSSPMDM_H */
