// libxsmm_b200 -- host helper shared by the kernels that use TMA tensor maps: resolves
// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda).
#ifndef XB_TMA_CUH
#define XB_TMA_CUH
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*xb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline xb_encode_tiled_fn xb_tma_encoder() {
  static xb_encode_tiled_fn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr) { (void)cudaGetLastError(); return nullptr; }
    fn = (xb_encode_tiled_fn)p;
  }
  return fn;
}
#endif
