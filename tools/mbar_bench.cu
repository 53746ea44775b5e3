// Micro-benchmark (diagnostic): round-trip latency of an mbarrier ping-pong between two warps of one CTA.
//  mode bit0: waiting side polls with test_wait (else try_wait);  bit1: only lane 0 of a warp polls/arrives, else all 32 lanes poll and
//  lane 0 arrives;  extra idle pollers (warps spinning on a third barrier that never completes) model a crowded CTA.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool poll(uint32_t bar, uint32_t parity, int test) {
  uint32_t done;
  if (test == 2) { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity), "r"(1000000u) : "memory"); return done != 0; }
  if (test == 3) { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory"); if (!done) __nanosleep(1000); return done != 0; }
  if (test) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  else asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__global__ void k(int mode, int reps, int idle_warps, int idle_mode, long long* out) {
  __shared__ uint64_t bars[4];
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[i]))); stop = 0; }
  __syncthreads();
  const uint32_t ping = smem_u32(&bars[0]), pong = smem_u32(&bars[1]), never = smem_u32(&bars[2]);
  const int test = (mode & 4) ? 2 : (mode & 1), lane_only = (mode >> 1) & 1;
  if (warp == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (lane == 0) arrive(ping);
      if (!lane_only || lane == 0) while (!poll(pong, r & 1, test)) {}
      __syncwarp();
    }
    if (lane == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
    if (lane == 0) stop = 1;
  } else if (warp == 1) {
    for (int r = 0; r < reps; ++r) {
      if (!lane_only || lane == 0) while (!poll(ping, r & 1, test)) {}
      __syncwarp();
      if (lane == 0) arrive(pong);
    }
  } else if (warp - 2 < idle_warps) {
    while (!stop) { if (idle_mode == 4) { if (lane == 0) poll(never, 0, 3); __syncwarp(); } else poll(never, 0, idle_mode); }
  }
}
int main() {
  long long* out; cudaMallocManaged(&out, 16);
  const int reps = 4000;
  const char* im[] = {"try_wait", "test_wait", "try_wait+hint(1ms)", "try_wait+nanosleep(1us)", "lane0 try_wait+nanosleep(1us)"};
  for (int idle_mode = 0; idle_mode < 5; ++idle_mode) for (int mode : {0, 2, 4}) {
    const int idle = 18;
    k<<<148, 32 * (2 + idle)>>>(mode, reps, idle, idle_mode, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("18 idle warps polling with %-30s | ping-pong: %s %s : %.1f cycles per round trip %s\n", im[idle_mode], (mode & 4) ? "try_wait+hint" : "try_wait     ", (mode & 2) ? "lane0 polls  " : "32 lanes poll",
           out[0] / (double)reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
