// Diagnostic: does a tiled tensor map with elementStrides[0] = 2 (INTERLEAVE_NONE) de-interleave 16-bit pairs?
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../libxsmm_b200/csrc/xb_tma.cuh"
__global__ void k(const __grid_constant__ CUtensorMap map, unsigned short* out, int start, int nbytes) {
  __shared__ __align__(1024) unsigned short buf[4096];
  __shared__ uint64_t bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(buf);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4096; ++i) buf[i] = 0xFFFF;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(nbytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(d), "l"(&map), "r"(start), "r"(0), "r"(b) : "memory");
    uint32_t done; int spins = 0;
    do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(0u) : "memory"); } while (!done && ++spins < 2000000);
    out[4096] = (unsigned short)done;
    for (int i = 0; i < 4096; ++i) out[i] = buf[i];
  }
}
int main(int argc, char** argv) {
  const int es_arg = argc > 1 ? atoi(argv[1]) : 2, start_arg = argc > 2 ? atoi(argv[2]) : 0, box_arg = argc > 3 ? atoi(argv[3]) : 128;
  xb_encode_tiled_fn enc = xb_tma_encoder();
  if (!enc) { printf("no encoder\n"); return 1; }
  unsigned short* g; cudaMallocManaged(&g, 64 * 128 * 2); unsigned short* out; cudaMallocManaged(&out, 4097 * 2);
  for (int r = 0; r < 64; ++r) for (int x = 0; x < 128; ++x) g[r * 128 + x] = (unsigned short)(r * 1000 + x);   // row r: 64 (m,t) pairs, value = 1000 r + 2 m + t
  for (int es = es_arg; es <= es_arg; ++es) for (int start = start_arg; start <= start_arg; ++start) for (int boxx : {box_arg}) {
    CUtensorMap map;
    const cuuint64_t dims[2] = {128, 64}; const cuuint64_t strides[1] = {256};
    const cuuint32_t box[2] = {(cuuint32_t)boxx, 4}; const cuuint32_t estr[2] = {(cuuint32_t)es, 1};
    CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, g, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("es=%d start=%d box=%d: encode failed rc=%d\n", es, start, boxx, (int)rc); continue; }
    for (int nb : {boxx * 4 * 2 / es, boxx * 4 * 2}) {
      for (int i = 0; i < 4097; ++i) out[i] = 0;
      k<<<1, 32>>>(map, out, start, nb);
      cudaError_t e = cudaDeviceSynchronize();
      printf("es=%d start=%d box=%d expect_tx=%d: %s done=%d | row0:", es, start, boxx, nb, cudaGetErrorString(e), out[4096]);
      for (int i = 0; i < 10; ++i) printf(" %d", out[i]);
      printf(" ... [%d]=%d [%d]=%d [%d]=%d\n", boxx / es - 1, out[boxx / es - 1], boxx / es, out[boxx / es], boxx, out[boxx]);
      if (e != cudaSuccess) return 0;
      if (es == 1) break;
    }
  }
  return 0;
}
