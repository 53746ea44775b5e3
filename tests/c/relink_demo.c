/* A plain C caller of the LIBXSMM API, written for this repository's tests: it is compiled against include/ and linked
 * with -lxsmm (the libxsmm.so symlink of libxsmm_b200.so), the way an existing LIBXSMM user would relink.
 *
 *   relink_demo dispatch   host-only checks: handles are non-NULL, identical descriptors give identical pointers,
 *                          kernel info is consistent (runs without a GPU)
 *   relink_demo run        additionally calls the handle 200 times on HOST buffers, C += A_i * B_i (13x5x7, F64 like the
 *                          reference's samples/hello), and compares with a plain triple loop: prints "max_abs_diff <x>"
 */
#include <libxsmm.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

int main(int argc, char* argv[]) {
  const libxsmm_blasint m = 13, n = 5, k = 7, batch = 200;
  const libxsmm_gemm_shape shape = libxsmm_create_gemm_shape(m, n, k, m, k, m, LIBXSMM_DATATYPE_F64, LIBXSMM_DATATYPE_F64,
                                                           LIBXSMM_DATATYPE_F64, LIBXSMM_DATATYPE_F64);
  const libxsmm_gemmfunction kernel = libxsmm_dispatch_gemm(shape, LIBXSMM_GEMM_FLAG_NONE, LIBXSMM_GEMM_PREFETCH_NONE);
  const libxsmm_gemmfunction again = libxsmm_dispatch_gemm(shape, LIBXSMM_GEMM_FLAG_NONE, LIBXSMM_GEMM_PREFETCH_NONE);
  const libxsmm_gemmfunction beta0 = libxsmm_dispatch_gemm(shape, LIBXSMM_GEMM_FLAG_BETA_0, LIBXSMM_GEMM_PREFETCH_NONE);
  libxsmm_kernel_info info;
  libxsmm_xmmfunction x;
  libxsmm_mmkernel_info mm;
  if (kernel == NULL || kernel != again || beta0 == NULL || beta0 == kernel) { printf("dispatch identity failed\n"); return 1; }
  if (0 != libxsmm_get_kernel_info((const void*)kernel, &info) || info.nflops != 2u * 13 * 5 * 7) { printf("kernel info failed\n"); return 1; }
  x.gemm = kernel;
  if (0 != libxsmm_get_mmkernel_info(x, &mm) || mm.m != 13 || mm.n != 5 || mm.k != 7) { printf("mmkernel info failed\n"); return 1; }
  printf("target %s, typesize(f64) %d\n", libxsmm_get_target_arch(), (int)libxsmm_typesize(LIBXSMM_DATATYPE_F64));
  if (argc > 1 && 0 == strcmp(argv[1], "run")) {
    double* a = (double*)malloc(sizeof(double) * batch * m * k);
    double* b = (double*)malloc(sizeof(double) * batch * k * n);
    double* c = (double*)calloc((size_t)m * n, sizeof(double));
    double* want = (double*)calloc((size_t)m * n, sizeof(double));
    libxsmm_gemm_param p;
    libxsmm_blasint t, i, j, s;
    double diff = 0;
    for (i = 0; i < batch * m * k; ++i) a[i] = (double)((i * 7) % 23 - 11) / 8.0;
    for (i = 0; i < batch * k * n; ++i) b[i] = (double)((i * 5) % 19 - 9) / 4.0;
    memset(&p, 0, sizeof(p));
    p.c.primary = c;
    for (t = 0; t < batch; ++t) {
      p.a.primary = a + t * m * k; p.b.primary = b + t * k * n;
      kernel(&p);                                                  /* C += A_t * B_t, exactly as a CPU caller would */
      for (j = 0; j < n; ++j) for (i = 0; i < m; ++i) for (s = 0; s < k; ++s) want[j * m + i] += a[t * m * k + s * m + i] * b[t * k * n + j * k + s];
    }
    for (i = 0; i < m * n; ++i) { const double e = fabs(c[i] - want[i]); if (e > diff) diff = e; }
    printf("max_abs_diff %.3e\n", diff);
    free(a); free(b); free(c); free(want);
    return (diff == 0) ? 0 : 2;     /* same operation order and no FMA contraction on either side: exact */
  }
  printf("dispatch ok\n");
  return 0;
}
