/* TEST HELPER: exposes macro-only parts of include/libxsmm_utils.h to ctypes (LIBXSMM_MATINIT, LIBXSMM_DATATYPE). */
#include <libxsmm.h>
#include <libxsmm_utils.h>

void probe_matinit(int is_f64, double seed, void* dst, int nrows, int ncols, int ld, double scale) {
  if (is_f64) { LIBXSMM_MATINIT(double, seed, dst, nrows, ncols, ld, scale); }
  else { LIBXSMM_MATINIT(float, seed, dst, nrows, ncols, ld, scale); }
}
int probe_datatype_double(void) { return (int)LIBXSMM_DATATYPE(double); }
int probe_datatype_float(void) { return (int)LIBXSMM_DATATYPE(float); }
int probe_flags(void) { return (int)LIBXSMM_GEMM_FLAGS('N', 'T') | (LIBXSMM_NEQ(0, 1.0) ? 1024 : 0); }
