// Micro-benchmark (diagnostic): issue cost / service time of tcgen05.mma kind::f16 for several shapes and operand majors.
// One CTA per SM, one elected lane issues REPS MMAs back to back on (garbage) shared memory, commit, wait, report cycles/MMA.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a, uint32_t lbo16, uint32_t sbo16, uint32_t layout) {
  return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)(lbo16 & 0x3FFFu) << 16) | ((uint64_t)(sbo16 & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
template <int M, int N, int a_mn, int ALT>
__global__ void __launch_bounds__(128, 1) k(int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar; __shared__ uint32_t tw;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tw)), "r"(512u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  asm volatile("fence.proxy.async.shared::cta;");
  const uint32_t tb = tw;
  if (threadIdx.x < 32) {
    uint32_t leader; asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t ad = a_mn ? make_desc(smem_u32(smem), 8192 >> 4, 1024 >> 4, 2) : make_desc(smem_u32(smem), 1, 1024 >> 4, 2);
    const uint64_t bd = make_desc(smem_u32(smem) + 32768, 1, 1024 >> 4, 2);
    const long long t0 = clock64();
#pragma unroll 4
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tb + (ALT ? (uint32_t)((r & 1) * N) : 0u);
      if (leader) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" :: "r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u));
    }
    const long long t1 = clock64();
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    uint32_t done; do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory"); } while (!done);
    const long long t2 = clock64();
    if (leader && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (threadIdx.x < 32) { asm volatile("tcgen05.fence::after_thread_sync;"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(512u)); }
}
template <int M, int N, int a_mn, int ALT> void run(long long* out) {
  const int reps = 2048;
  cudaFuncSetAttribute(k<M, N, a_mn, ALT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<M, N, a_mn, ALT><<<148, 128, 100 * 1024>>>(reps, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("M=%3d N=%3d A=%s alt=%d : issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %.0f) %s\n", M, N, a_mn ? "MN" : "K ", ALT, out[0] / (double)reps, out[1] / (double)reps,
         (M > 128 ? M : 128) * N / 256.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
}
template <int M, int a_mn> void runN(long long* out) {
  run<M, 16, a_mn, 0>(out); run<M, 32, a_mn, 0>(out); run<M, 32, a_mn, 1>(out); run<M, 64, a_mn, 0>(out); run<M, 64, a_mn, 1>(out);
  run<M, 128, a_mn, 0>(out); run<M, 128, a_mn, 1>(out); run<M, 256, a_mn, 0>(out); run<M, 256, a_mn, 1>(out);
}
int main() {
  long long* out; cudaMallocManaged(&out, 16);
  runN<64, 0>(out); runN<64, 1>(out); runN<128, 0>(out); runN<128, 1>(out);
  return 0;
}
