// libxsmm_b200 -- thin CUDA runtime shim used by the plain-C host code: device/stream selection per
// host thread, sticky error reporting (kernels return void, like the reference's handles), memory,
// pointer classification and a per-thread device scratch arena used to stage host-resident operands.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include "xb_internal.h"

namespace {
struct ThreadState {
  cudaStream_t stream = nullptr;
  int blocking = 1;
  char* scratch = nullptr; size_t scratch_cap = 0, scratch_used = 0;
  // overflow blocks (kept until reset) when a request does not fit the arena
  void* spill[64]; int nspill = 0;
};
thread_local ThreadState tls;
std::atomic<int> g_last_error{0};
std::atomic<unsigned long long> g_launches{0};
char g_error_where[128] = {0};
int g_have_gpu = -1;
}  // namespace

extern "C" {

int xb_rt_have_gpu(void) {
  if (g_have_gpu < 0) {
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    g_have_gpu = (e == cudaSuccess && n > 0) ? 1 : 0;
    if (e != cudaSuccess) (void)cudaGetLastError();
  }
  return g_have_gpu;
}

int xb_rt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  return n;
}

int xb_rt_set_device(int ordinal) {
  const cudaError_t e = cudaSetDevice(ordinal);
  if (e != cudaSuccess) { xb_rt_note_error((int)e, "set_device"); return (int)e; }
  return 0;
}

void xb_rt_set_stream(void* stream) { tls.stream = (cudaStream_t)stream; }
void* xb_rt_stream(void) { return (void*)tls.stream; }
void xb_rt_set_blocking(int on) { tls.blocking = on ? 1 : 0; }
int xb_rt_blocking(void) { return tls.blocking; }

void xb_rt_note_error(int code, const char* where) {
  int expected = 0;
  bool first = false;
  if (g_last_error.compare_exchange_strong(expected, code)) {
    snprintf(g_error_where, sizeof(g_error_where), "%s", where ? where : "?");
    first = true;
  }
  // handles return void (reference ABI): the first failure is always reported, there is no CPU path to fall back to
  if (first || libxsmm_verbosity != 0) {
    fprintf(stderr, "LIBXSMM-B200 ERROR (%s): %s\n", where ? where : "?", cudaGetErrorString((cudaError_t)code));
  }
}

int xb_rt_sync(void) {
  const cudaError_t e = cudaStreamSynchronize(tls.stream);
  if (e != cudaSuccess) xb_rt_note_error((int)e, "sync");
  return g_last_error.load();
}

int xb_rt_last_error(void) { return g_last_error.load(); }
const char* xb_rt_last_error_string(void) {
  static thread_local char buf[256];
  const int e = g_last_error.load();
  if (e == 0) return "";
  snprintf(buf, sizeof(buf), "%s: %s", g_error_where, cudaGetErrorString((cudaError_t)e));
  return buf;
}

unsigned long long xb_rt_launch_count(void) { return g_launches.load(); }
void xb_rt_count_launch(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }

void* xb_rt_device_malloc(size_t size) {
  void* p = nullptr;
  const cudaError_t e = cudaMalloc(&p, size ? size : 1);
  if (e != cudaSuccess) { xb_rt_note_error((int)e, "device_malloc"); return nullptr; }
  return p;
}
void xb_rt_device_free(void* p) { if (p) cudaFree(p); }
void* xb_rt_host_malloc(size_t size) {
  void* p = nullptr;
  const cudaError_t e = cudaMallocHost(&p, size ? size : 1);
  if (e != cudaSuccess) { xb_rt_note_error((int)e, "host_malloc"); return nullptr; }
  return p;
}
void xb_rt_host_free(void* p) { if (p) cudaFreeHost(p); }
void* xb_rt_managed_malloc(size_t size) {
  void* p = nullptr;
  const cudaError_t e = cudaMallocManaged(&p, size ? size : 1, cudaMemAttachGlobal);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
  return p;
}
void xb_rt_managed_free(void* p) { if (p) cudaFree(p); }

int xb_rt_memcpy(void* dst, const void* src, size_t size) {
  if (size == 0) return 0;
  cudaError_t e = cudaMemcpyAsync(dst, src, size, cudaMemcpyDefault, tls.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(tls.stream);
  if (e != cudaSuccess) { xb_rt_note_error((int)e, "memcpy"); return (int)e; }
  return 0;
}
int xb_rt_memcpy_async(void* dst, const void* src, size_t size) {
  if (size == 0) return 0;
  const cudaError_t e = cudaMemcpyAsync(dst, src, size, cudaMemcpyDefault, tls.stream);
  if (e != cudaSuccess) { xb_rt_note_error((int)e, "memcpy_async"); return (int)e; }
  return 0;
}
int xb_rt_upload(void* dst_dev, const void* src_host, size_t size) { return xb_rt_memcpy_async(dst_dev, src_host, size); }

int xb_rt_ptr_kind(const void* p) {
  if (p == nullptr) return 0;
  cudaPointerAttributes at;
  const cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  switch (at.type) {
    case cudaMemoryTypeDevice: return 1;
    case cudaMemoryTypeManaged: return 2;
    case cudaMemoryTypeHost: return 3;
    default: return 0;
  }
}

void* xb_rt_scratch(size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;
  if (tls.scratch_used + bytes <= tls.scratch_cap) {
    void* p = tls.scratch + tls.scratch_used; tls.scratch_used += bytes; return p;
  }
  if (tls.scratch_used == 0) {               // nothing handed out: grow the arena
    if (tls.scratch) cudaFree(tls.scratch);
    size_t cap = tls.scratch_cap ? tls.scratch_cap : ((size_t)1 << 22);
    while (cap < bytes) cap <<= 1;
    tls.scratch = nullptr; tls.scratch_cap = 0;
    if (cudaMalloc((void**)&tls.scratch, cap) != cudaSuccess) { xb_rt_note_error((int)cudaGetLastError(), "scratch"); return nullptr; }
    tls.scratch_cap = cap; tls.scratch_used = bytes; return tls.scratch;
  }
  if (tls.nspill < 64) {                     // arena in use: one-off block until the next reset
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { xb_rt_note_error((int)cudaGetLastError(), "scratch"); return nullptr; }
    tls.spill[tls.nspill++] = p; return p;
  }
  return nullptr;
}

void xb_rt_scratch_reset(void) {
  for (int i = 0; i < tls.nspill; ++i) cudaFree(tls.spill[i]);
  tls.nspill = 0; tls.scratch_used = 0;
}

}  // extern "C"
