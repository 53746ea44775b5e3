// libxsmm_b200 -- packed block-sparse (BCSC) B x dense A on the tcgen05 tensor cores (sm_100a), bf16.
//
// For every m_block mb:  C_mb[N][M] = beta * C_mb + A_mb(M x K) * B(K x N),  B given as BCSC blocks [bn][bk].
// Replaces src/generator_packed_spgemm_bcsc_bsparse_avx_avx2_avx512_amx.c (AMX/AVX-512 per block);
// semantics are those of the driver's dense gold (samples/xgemm_sparse/spmm_kernel.c:74-217) with f32
// accumulation in tensor memory (tolerance 5e-3 for bf16 like spmm_kernel.c:1019-1029).
//
// Mapping. B is shared by all m_blocks, so 128/M consecutive m_blocks are stacked into ONE M=128 MMA operand
// ("group"). A work item is (group, column part): D(128 x <=256 columns) lives in TMEM, which therefore holds
// two items and lets the epilogue of one overlap the MMAs of the next. Every non-zero block (kb, j) contributes
// bk/16 instructions  D[:, j*bn : (j+1)*bn] += A_grp[:, kb*bk : (kb+1)*bk] * B_blk^T;  adjacent blocks of a
// block-row are merged into one wider instruction. Blocks are visited in K-MAJOR order, 64 k at a time
// ("k-step" = 64/bk block-rows), so that A streams through a small ring while the accumulator columns stay
// resident. Two tiny kernels run in front on every call (pattern and values arrive with the call):
//   bcsc_prep_kernel   re-sorts the CSC block pattern into per (part, k-step) lists of ready-to-issue MMA
//                      operations; it first compares the pattern with the copy cached in the handle and returns
//                      at once when nothing changed
//   bcsc_pack_b_kernel copies the B blocks into visiting order, pre-swizzled to the shared-memory image the MMA
//                      descriptor expects, so that one k-step of a part is ONE contiguous bulk copy
//
// Pipeline per CTA (persistent, one per SM, 22 warps):
//   warp 0        A producer: TMA 3D box of the raw VNNI A words of one k-step (128 rows x 64 k = 16 KB)
//   warp 2        B producer: one cp.async.bulk per k-step (the packed blocks of that list)
//   warps 12-19   converters: VNNI2 words -> two k-rows of the canonical MN-major SWIZZLE_128B operand
//                             (16-bit de-interleave with PRMT, 16-byte shared stores), fence.proxy.async
//   warps 1,3,20,21 MMA issuers: tcgen05.mma.cta_group::1.kind::f16, M=128, N=run*bn, K=16; each warp OWNS a range
//                             of block-columns, so every accumulator column is written by one thread in k order
//   warps 4-11    epilogue  : tcgen05.ld -> bf16 (hardware RNE) -> packed 4-byte stores
// Measured design notes (profiles/): a stage hand-over costs ~600 cycles of mbarrier round trips whatever the
// payload, hence the 64-wide k-steps; 2D tensor-map loads of [bn][bk] blocks cost the TMA unit ~5 cycles per
// 64-byte row, hence the repack; tcgen05.mma (M=128, K=16) costs max(N/2, 32 + N/4) cycles, A and B major alike.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "xb_internal.h"
#include "xb_device.cuh"
#include "xb_tma.cuh"

namespace {

struct BcscTcParams {
  int bn, bk, nbc, nks, ksteps;             // nks: 64-wide k-steps, ksteps = bk / 16
  long long m_blocks, ngroups;
  int ncols, slot_cols, nslot;              // N = nbc*bn, TMEM columns per work item, slots
  int nparts, part_cols;                    // column parts per group (each <= 256 columns so that TMEM holds >= 2 items)
  int raw_stages, can_stages, b_stages, b_stage_bytes;
  const unsigned int* list_ptr;             // [nparts*nks + 1] first block of every (part, k-step) list
  const unsigned int* wranges;              // [4 * lists] {first op | count << 16} per MMA warp
  const uint4* ops;                         // {d col | A row offset/16 << 16, B offset/16, idesc N bits, accumulate}
  const unsigned int* col_any;              // [nbc] column has at least one block
  const char* b_packed;                     // B blocks in visiting order, pre-swizzled (bcsc_pack_b_kernel)
  int mma_warps, ops_cap;                   // ops_cap: operations that fit the shared-memory copy (variant 2)
  char* c; int beta0;
  uint32_t idesc, b_layout, b_sbo16;        // UMMA descriptor pieces
  int spin;                                 // bit mask: roles polling with test_wait (1 MMA, 2 epilogue, 4 converters, 8 producers)
  int b_cpasync;                            // experiment (LIBXSMM_B200_BCSC_BCPASYNC=1): B through per-thread cp.async instead of the TMA engine
  int sleep;                                // producers/converters/epilogue back off with nanosleep between polls (LIBXSMM_B200_BCSC_SLEEP)
  int skip;                                 // diagnostic ablation mask (LIBXSMM_B200_BCSC_SKIP): 1 conv, 2 B loads, 4 A loads, 8 stores, 16 MMAs
  long long* dbg;                           // optional per-role cycle counters (CTA 0), see tools/bcsc_probe.py
};

// layout of the per-handle index buffer (words); shared by host and prep kernel
struct BcscIdxLayout {
  unsigned int hdr, cache_cp, cache_ri, list_ptr, entries, col_any, wranges, ops, total;
  __host__ __device__ BcscIdxLayout(unsigned int nbc, unsigned int nl, unsigned int cap) {
    hdr = 0; cache_cp = 16; cache_ri = cache_cp + nbc + 1; list_ptr = cache_ri + cap; entries = list_ptr + nl + 1;
    col_any = entries + cap; wranges = col_any + nbc; ops = (wranges + 4 * nl + 3) & ~3u; total = ops + 4 * cap + 4;
  }
};


__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
// for waits that are expected to be long: back off so that polling warps leave issue slots to the working ones
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, unsigned int ns) {
  uint32_t done;
  for (;;) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (ns != 0) __nanosleep(ns);
  }
}
// spin != 0: poll with the non-blocking test_wait instead of the (hardware-suspending) try_wait
__device__ __forceinline__ void mbar_wait_x(uint32_t bar, uint32_t parity, int spin) {
  if (!spin) { mbar_wait(bar, parity); return; }
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// high word of an SMEM matrix descriptor (stride offset, version 1, layout); the low word carries address and LBO
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo16, uint32_t layout) { return (sbo16 & 0x3FFFu) | (1u << 14) | (layout << 29); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo16) { return ((smem_addr & 0x3FFFFu) >> 4) | ((lbo16 & 0x3FFFu) << 16); }
__device__ __forceinline__ uint64_t desc64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

// ---- prep: CSC block pattern -> per (part, k-step) lists of MMA operations ---------------------------------------------
// Every MMA warp OWNS cpw adjacent block-columns of a part (merged runs never cross owners): all instructions on one
// accumulator column are issued by one thread in k order, so 'overwrite' always precedes 'accumulate' and the summation
// order is fixed without any cross-warp synchronisation. Lists are stored owner-major: (warp, block-row, column).
__global__ void __launch_bounds__(256) bcsc_prep_kernel(const unsigned int* __restrict__ colptr, const unsigned int* __restrict__ rowidx,
                                                        int nbc, int nkb, int bn, int bk, int nparts, int bpp, int cpw, unsigned int cap, unsigned int* buf) {
  const int KBS = 64 / bk, nks = (nkb + KBS - 1) / KBS, nl = nparts * nks;
  const BcscIdxLayout L((unsigned int)nbc, (unsigned int)nl, cap);
  __shared__ unsigned int s_cp[257];
  __shared__ unsigned int s_ri[2048];
  __shared__ unsigned short s_map[2048];          // dense (block-row, column) -> block index + 1
  __shared__ unsigned int cnt[513];
  __shared__ unsigned int s_kmin[256];
  for (int j = threadIdx.x; j <= nbc; j += blockDim.x) s_cp[j] = colptr[j];
  __syncthreads();
  unsigned int nnzb = s_cp[nbc]; if (nnzb > cap) nnzb = cap;
  for (unsigned int z = threadIdx.x; z < nnzb; z += blockDim.x) s_ri[z] = rowidx[z];
  // unchanged pattern (the usual case: weights are re-used across calls)? compare in full with the cached copy
  const unsigned int key[9] = {0xb200c5c5u, (unsigned int)nbc, (unsigned int)nkb, (unsigned int)bn, (unsigned int)bk, (unsigned int)nparts,
                               (unsigned int)bpp, (unsigned int)cpw, nnzb};
  int same = 1;
  if (threadIdx.x < 9) same = (buf[L.hdr + threadIdx.x] == key[threadIdx.x]);
  for (int j = threadIdx.x; j <= nbc; j += blockDim.x) same &= (buf[L.cache_cp + j] == s_cp[j]);
  for (unsigned int z = threadIdx.x; z < nnzb; z += blockDim.x) same &= (buf[L.cache_ri + z] == s_ri[z]);
  if (__syncthreads_and(same)) return;

  for (int i = threadIdx.x; i <= nl; i += blockDim.x) cnt[i] = 0;
  for (int i = threadIdx.x; i < nbc * nkb; i += blockDim.x) s_map[i] = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < nbc; j += blockDim.x) {
    unsigned int kmin = 0xffffffffu;
    for (unsigned int z = s_cp[j]; z < s_cp[j + 1] && z < nnzb; ++z) {
      const unsigned int kb = s_ri[z];
      if (kb < (unsigned int)nkb) {
        s_map[kb * nbc + j] = (unsigned short)(z + 1);
        atomicAdd(&cnt[(j / bpp) * nks + kb / KBS + 1], 1u);
        kmin = (kb < kmin) ? kb : kmin;
      }
    }
    s_kmin[j] = kmin; buf[L.col_any + j] = (kmin != 0xffffffffu) ? 1u : 0u;
  }
  __syncthreads();
  if (threadIdx.x == 0) { for (int i = 0; i < nl; ++i) cnt[i + 1] += cnt[i]; }
  __syncthreads();
  for (int i = threadIdx.x; i <= nl; i += blockDim.x) buf[L.list_ptr + i] = cnt[i];
  uint4* ops = reinterpret_cast<uint4*>(buf + L.ops);
  const int maxrun = 256 / bn;
  const unsigned int blk16 = (unsigned int)(bn * bk * 2) >> 4;
  for (int l = threadIdx.x; l < nl; l += blockDim.x) {
    const int part = l / nks, ks = l % nks, j0 = part * bpp, j1 = (j0 + bpp < nbc) ? j0 + bpp : nbc;
    unsigned int e = cnt[l], o = cnt[l], pos = 0;
    for (int w = 0; w < 4; ++w) {
      const int c0 = j0 + w * cpw, c1 = (c0 + cpw < j1) ? c0 + cpw : j1;
      const unsigned int ostart = o;
      for (int kbi = 0; kbi < KBS && ks * KBS + kbi < nkb; ++kbi) {
        const int kb = ks * KBS + kbi;
        const unsigned int a_off16 = (unsigned int)(kbi * bk * 128) >> 4;
        int run_j = 0, run_len = 0; unsigned int run_first = 0, run_pos = 0;
        for (int j = c0; j < c1; ++j) {
          const unsigned int zz = s_map[kb * nbc + j];
          if (zz == 0) continue;
          const unsigned int first = ((unsigned int)kb == s_kmin[j]) ? 1u : 0u;
          buf[L.entries + e++] = zz - 1;
          if (run_len > 0 && j == run_j + run_len && first == run_first && run_len < maxrun) ++run_len;
          else {
            if (run_len > 0) ops[o++] = make_uint4((unsigned int)((run_j - j0) * bn) | (a_off16 << 16), run_pos * blk16, ((unsigned int)(run_len * bn) >> 3) << 17, run_first ^ 1u);
            run_j = j; run_len = 1; run_first = first; run_pos = pos;
          }
          ++pos;
        }
        if (run_len > 0) ops[o++] = make_uint4((unsigned int)((run_j - j0) * bn) | (a_off16 << 16), run_pos * blk16, ((unsigned int)(run_len * bn) >> 3) << 17, run_first ^ 1u);
      }
      buf[L.wranges + 4 * l + w] = ostart | ((o - ostart) << 16);
    }
  }
  for (int j = threadIdx.x; j <= nbc; j += blockDim.x) buf[L.cache_cp + j] = s_cp[j];
  for (unsigned int z = threadIdx.x; z < nnzb; z += blockDim.x) buf[L.cache_ri + z] = s_ri[z];
  if (threadIdx.x < 9) buf[L.hdr + threadIdx.x] = key[threadIdx.x];
}

// ---- B repack: blocks in visiting order, each already in the shared-memory image the MMA descriptor expects ------------
// (K-major, rows of S = 2*bk bytes, 16-byte chunks XOR-swizzled exactly like the hardware SWIZZLE_<S>B mode). One k-step
// of a column part is then ONE contiguous run and is fetched with a single bulk copy. (Fetching the caller's blocks with a
// 2D tensor map costs the TMA unit ~5 cycles per 64-byte row: measured 287K of 357K cycles per CTA.)
__global__ void __launch_bounds__(256) bcsc_pack_b_kernel(const uint4* __restrict__ b_vals, const unsigned int* __restrict__ entries,
                                                          const unsigned int* __restrict__ list_ptr, int nl, uint4* __restrict__ out, int bn, int bk) {
  const unsigned int nnzb = list_ptr[nl];
  const unsigned int cpr = (unsigned int)bk >> 3;                    // 16-byte chunks per row (2, 4 or 8)
  const unsigned int cpb = (unsigned int)bn * cpr;                   // chunks per block
  const unsigned long long total = (unsigned long long)nnzb * cpb;
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned int e = (unsigned int)(t / cpb), ci = (unsigned int)(t % cpb);
    const unsigned int n = ci / cpr, c = ci % cpr;
    const unsigned int z = entries[e];
    const unsigned int sw = ((n * cpr * 16u) >> 7) & (cpr - 1u);     // Swizzle<log2(cpr),4,3>: address bits [7,..) folded onto the chunk bits
    out[(size_t)e * cpb + n * cpr + (c ^ sw)] = b_vals[(size_t)z * cpb + ci];
  }
}

// cycle accounting per role is compiled only into the DBG instantiation: clock64() reads around every wait cost more than the waits
#define XB_TWAIT(acc, call) do { call; } while (0)
// DBG: count the waits that found their barrier already complete (a role that never has to wait is the bottleneck)
#define XB_READY(acc, bar, parity) do { if (DBG) { uint32_t d_; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(d_) : "r"(bar), "r"(parity) : "memory"); acc += d_; } } while (0)
#define XB_CLOCK() (DBG ? clock64() : 0ll)

// ---- main kernel -------------------------------------------------------------------------------------------------------
constexpr int kThreads = 704, kConvWarps = 8;      // 22 warps: A prod, B prod, 4 MMA, 8 epilogue, 8 converters
constexpr int kMaxStages = 16, kMaxSlots = 4, kMaxEntries = 2048, kMaxLists = 512;
constexpr int A_STAGE = 128 * 64 * 2;               // bytes of one k-step of A (raw or canonical): 128 rows x 64 k

template <int M, bool DBG>
__global__ void __launch_bounds__(kThreads, 1)
bcsc_tc_kernel(const __grid_constant__ CUtensorMap map_a, const BcscTcParams P) {
  constexpr int G = 128 / M;                       // m_blocks per group
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // carve: canonical A ring | raw A ring | B ring | operations | list pointers | per-warp ranges | col flags | barriers | tmem word
  uint8_t* s_can = smem;
  uint8_t* s_raw = s_can + (size_t)P.can_stages * A_STAGE;
  uint8_t* s_b = s_raw + (size_t)P.raw_stages * A_STAGE;
  uint4* s_ops = (uint4*)(s_b + (size_t)P.b_stages * P.b_stage_bytes);
  unsigned int* s_lp = (unsigned int*)(s_ops + kMaxEntries + 1);       // +1: the MMA loop reads one operation ahead
  unsigned int* s_wr = s_lp + kMaxLists + 8;
  unsigned char* s_any = (unsigned char*)(s_wr + 4 * kMaxLists);
  uint64_t* bars = (uint64_t*)(s_any + 1024);
  const uint32_t bar0 = smem_u32(bars);
  const int RS = P.raw_stages, CS = P.can_stages, BS = P.b_stages, NS = P.nslot;
  const uint32_t raw_full = bar0, raw_empty = raw_full + 8 * kMaxStages, can_full = raw_empty + 8 * kMaxStages, can_empty = can_full + 8 * kMaxStages;
  const uint32_t b_full = can_empty + 8 * kMaxStages, b_empty = b_full + 8 * kMaxStages;
  const uint32_t t_full = b_empty + 8 * kMaxStages, t_empty = t_full + 8 * kMaxSlots;
  uint32_t* tmem_word = (uint32_t*)(bars + 6 * kMaxStages + 2 * kMaxSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long Gd = gridDim.x, bid = blockIdx.x;
  const int NP = P.nparts, NKS = P.nks, NL = P.nparts * P.nks;
  // work items of this CTA: its groups, each as NP consecutive column parts (the second read of A hits L2)
  const long long n_local = ((bid < P.ngroups) ? (P.ngroups - bid + Gd - 1) / Gd : 0) * NP;
  const uint32_t blk_bytes = (uint32_t)P.bn * P.bk * 2;

  {
    const unsigned int nnzb = P.list_ptr[NL];
    for (unsigned int i = threadIdx.x; i < nnzb; i += blockDim.x) s_ops[i] = P.ops[i];      // #operations <= #blocks
    for (int i = threadIdx.x; i < 4 * NL; i += blockDim.x) s_wr[i] = P.wranges[i];
    for (int i = threadIdx.x; i <= NL; i += blockDim.x) s_lp[i] = P.list_ptr[i];
    for (int i = threadIdx.x; i < P.nbc; i += blockDim.x) s_any[i] = (unsigned char)P.col_any[i];
  }
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_a) : "memory");
    for (int i = 0; i < RS; ++i) { mbar_init(raw_full + 8 * i, 1); mbar_init(raw_empty + 8 * i, kConvWarps); }
    for (int i = 0; i < CS; ++i) { mbar_init(can_full + 8 * i, kConvWarps); mbar_init(can_empty + 8 * i, (uint32_t)P.mma_warps); }
    for (int i = 0; i < BS; ++i) { mbar_init(b_full + 8 * i, P.b_cpasync ? 32u : 1u); mbar_init(b_empty + 8 * i, (uint32_t)P.mma_warps); }
    for (int i = 0; i < NS; ++i) { mbar_init(t_full + 8 * i, (uint32_t)P.mma_warps); mbar_init(t_empty + 8 * i, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_word)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_word;

  if (warp == 0) {
    // ========================================= A producer =======================================
    if (lane == 0) {
      int rs = 0; uint32_t rph = 0; long long w0 = 0; const long long tstart = XB_CLOCK();
      for (long long i = 0; i < n_local; ++i) {
        const long long grp = bid + (i / NP) * Gd;
        for (int ks = 0; ks < NKS; ++ks) {
          XB_READY(w0, raw_empty + 8 * rs, rph ^ 1); XB_TWAIT(w0, mbar_wait_x(raw_empty + 8 * rs, rph ^ 1, P.spin & 8));
          mbar_expect_tx(raw_full + 8 * rs, (P.skip & 4) ? 0u : (uint32_t)A_STAGE);     // rows beyond K are zero-filled by the TMA unit
          if (!(P.skip & 4)) asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                       :: "r"(smem_u32(s_raw + (size_t)rs * A_STAGE)), "l"(&map_a), "r"(0), "r"(ks * 32), "r"((int)(grp * G)),
                          "r"(raw_full + 8 * rs) : "memory");
          if (++rs == RS) { rs = 0; rph ^= 1; }
        }
      }
      if (DBG && P.dbg != nullptr && bid == 0) { P.dbg[0] = w0; P.dbg[1] = XB_CLOCK() - tstart; }
    }
  } else if (warp == 2) {
    // ========================================= B producer =======================================
    if (P.b_cpasync) {
      // experiment: B is L2-resident, so it can bypass the TMA engine: every lane copies 16-byte chunks with cp.async and
      // arrives on the stage barrier when its own copies have landed (cp.async.mbarrier.arrive.noinc)
      int bs = 0; uint32_t bph = 0;
      for (long long i = 0; i < n_local; ++i) {
        const int l0 = (int)(i % NP) * NKS;
        for (int ks = 0; ks < NKS; ++ks) {
          const unsigned int e0 = s_lp[l0 + ks], e1 = s_lp[l0 + ks + 1];
          mbar_wait(b_empty + 8 * bs, bph ^ 1);
          const uint32_t dst = smem_u32(s_b + (size_t)bs * P.b_stage_bytes);
          const char* src = P.b_packed + (size_t)e0 * blk_bytes;
          const unsigned int nchunk = ((e1 - e0) * blk_bytes) >> 4;
          for (unsigned int c = (unsigned int)lane; c < nchunk; c += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + c * 16u), "l"(src + (size_t)c * 16) : "memory");
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(b_full + 8 * bs) : "memory");
          if (++bs == BS) { bs = 0; bph ^= 1; }
        }
      }
    } else if (lane == 0) {
      int bs = 0; uint32_t bph = 0; long long w0 = 0; const long long tstart = XB_CLOCK();
      for (long long i = 0; i < n_local; ++i) {
        const int l0 = (int)(i % NP) * NKS;
        for (int ks = 0; ks < NKS; ++ks) {
          const unsigned int e0 = s_lp[l0 + ks], e1 = s_lp[l0 + ks + 1];
          XB_READY(w0, b_empty + 8 * bs, bph ^ 1); XB_TWAIT(w0, mbar_wait_x(b_empty + 8 * bs, bph ^ 1, P.spin & 8));
          mbar_expect_tx(b_full + 8 * bs, (P.skip & 2) ? 0u : (e1 - e0) * blk_bytes);     // zero blocks: the barrier completes at once
          if (e1 > e0 && !(P.skip & 2)) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(s_b + (size_t)bs * P.b_stage_bytes)), "l"(P.b_packed + (size_t)e0 * blk_bytes), "r"((e1 - e0) * blk_bytes),
                            "r"(b_full + 8 * bs) : "memory");
          }
          if (++bs == BS) { bs = 0; bph ^= 1; }
        }
      }
      if (DBG && P.dbg != nullptr && bid == 0) { P.dbg[2] = w0; P.dbg[3] = XB_CLOCK() - tstart; }
    }
  } else if (warp == 1 || (warp == 3 && P.mma_warps >= 2) || (warp >= 20 && warp - 18 < P.mma_warps)) {
    // ========================================= MMA issuers ======================================
    // The whole warp runs the loop (warp-uniform control flow keeps descriptor arithmetic on the uniform datapath);
    // one elected lane issues the tcgen05 instructions.
    {
      uint32_t leader;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
      const int mw = (warp == 1) ? 0 : ((warp == 3) ? 1 : warp - 18);
      int cs = 0, bs = 0; uint32_t cph = 0, bph = 0; long long w_t = 0, w_c = 0, w_b = 0, w_i = 0; const long long tstart = XB_CLOCK();
      const uint32_t a_hi = desc_hi(1024 >> 4, 2), b_hi = desc_hi(P.b_sbo16, P.b_layout);
      const uint32_t a_lo0 = desc_lo(smem_u32(s_can), (uint32_t)(64 * 128) >> 4);   // MN-major SW128: LBO = distance of the two 64-row atoms
      const uint32_t b_lo0 = desc_lo(smem_u32(s_b), 1);
      const int ksteps = P.ksteps;
      for (long long i = 0; i < n_local; ++i) {
        const int slot = (int)(i % NS);
        XB_READY(w_t, t_empty + 8 * slot, (uint32_t)(((i / NS) & 1) ^ 1)); XB_TWAIT(w_t, mbar_wait_x(t_empty + 8 * slot, (uint32_t)(((i / NS) & 1) ^ 1), P.spin & 1));
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)(slot * P.slot_cols);
        const int l0 = (int)(i % NP) * NKS;
        for (int ks = 0; ks < NKS; ++ks) {
          const unsigned int rng = s_wr[4 * (l0 + ks) + mw], ob = rng & 0xFFFFu, on = rng >> 16;   // this warp's operations of the k-step
          uint4 op = s_ops[ob];                                          // fetched ahead of the waits (a stale/unused slot is harmless)
          XB_READY(w_c, can_full + 8 * cs, cph); XB_TWAIT(w_c, mbar_wait_x(can_full + 8 * cs, cph, P.spin & 1));
          XB_READY(w_b, b_full + 8 * bs, bph); XB_TWAIT(w_b, mbar_wait_x(b_full + 8 * bs, bph, P.spin & 1));
          if (P.b_cpasync) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async wrote through the generic proxy
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)cs * (A_STAGE >> 4);
          const uint32_t b_stage_lo = b_lo0 + (uint32_t)bs * ((uint32_t)P.b_stage_bytes >> 4);
          for (unsigned int o = 0; o < on; ++o) {
            const uint4 nxt = s_ops[ob + o + 1];
            const uint32_t d = d_base + (op.x & 0xFFFFu), a_op = a_lo + (op.x >> 16), b_lo = b_stage_lo + op.y, idesc = P.idesc | op.z;
            uint32_t accumulate = op.w;
            for (int kk = 0; kk < ksteps; ++kk) {
              if (leader && !(P.skip & 16)) umma_f16(d, desc64(a_hi, a_op + kk * (2048 >> 4)), desc64(b_hi, b_lo + kk * (32 >> 4)), idesc, accumulate);
              accumulate = 1;
            }
            op = nxt;
          }
          if (leader) {
            if (P.skip & 32) { mbar_arrive(can_empty + 8 * cs); mbar_arrive(b_empty + 8 * bs); }   // diagnostic (only valid without MMAs)
            else { umma_commit(can_empty + 8 * cs); umma_commit(b_empty + 8 * bs); }
          }
          __syncwarp();
          if (++cs == CS) { cs = 0; cph ^= 1; }
          if (++bs == BS) { bs = 0; bph ^= 1; }
        }
        if (leader) umma_commit(t_full + 8 * slot);
        __syncwarp();
      }
      if (DBG && P.dbg != nullptr && bid == 0 && leader && mw == 0) { P.dbg[4] = w_t; P.dbg[5] = w_c; P.dbg[6] = w_b; P.dbg[7] = XB_CLOCK() - tstart; P.dbg[13] = w_i; }
    }
  } else if (warp >= 4 && warp < 12) {
    // ========================================= epilogue (8 warps: two per TMEM quadrant, half the columns each) =====
    const int q = warp & 3, half = (warp - 4) >> 2;
    const int row = 32 * q + lane;                       // row of the 128-row group
    const int mbl = row / M, m = row % M;
    const bool bn32 = (P.bn % 32) == 0;
    long long w0 = 0; const long long tstart = XB_CLOCK();
    for (long long i = 0; i < n_local; ++i) {
      const int slot = (int)(i % NS);
      const long long grp = bid + (i / NP) * Gd;
      const long long mb = grp * G + mbl;
      const bool valid = mb < P.m_blocks;
      const int part = (int)(i % NP), pc0 = part * P.part_cols;                      // first column of this part
      const int pcols = (P.ncols - pc0 < P.part_cols) ? (P.ncols - pc0) : P.part_cols;
      const int nchunks = (pcols + 31) / 32, cbeg = half ? (nchunks + 1) / 2 : 0, cend = half ? nchunks : (nchunks + 1) / 2;
      __nv_bfloat16* cblk = reinterpret_cast<__nv_bfloat16*>(P.c) + ((size_t)mb * P.ncols + pc0) * M + m;
      XB_READY(w0, t_full + 8 * slot, (uint32_t)((i / NS) & 1)); XB_TWAIT(w0, mbar_wait_x(t_full + 8 * slot, (uint32_t)((i / NS) & 1), P.spin & 2));
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(slot * P.slot_cols) + ((uint32_t)(q * 32) << 16);
      if (cbeg == cend) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(t_empty + 8 * slot); }
      for (int ch = cbeg; ch < cend; ++ch) {
        const int c0 = ch * 32;
        uint32_t v[32];
        if (P.skip & 64) { if (ch + 1 == cend) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(t_empty + 8 * slot); } continue; }
        tmem_ld32(taddr + (uint32_t)c0, v);
        if (ch + 1 == cend) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(t_empty + 8 * slot); }
        {
          __nv_bfloat16* dst = cblk + (size_t)c0 * M;
          const int ncol = (pcols - c0 < 32) ? (pcols - c0) : 32;              // 16 or 32 (bn is a multiple of 16)
          const int jb = (pc0 + c0) / P.bn;
          const bool any0 = s_any[jb] != 0, any1 = bn32 ? any0 : (s_any[(pc0 + c0 + 16) / P.bn < P.nbc ? (pc0 + c0 + 16) / P.bn : jb] != 0);
          if (P.beta0 && (M % 2) == 0) {
            // lanes 2i/2i+1 hold rows m/m+1: after one exchange per column PAIR the even lane owns (m, m+1) of column c and the
            // odd lane (m-1, m) of column c+1, i.e. one 4-byte store per lane covers two elements (all lanes shuffle)
            const bool odd = lane & 1;
            unsigned int* dst32 = reinterpret_cast<unsigned int*>(dst - (odd ? 1 : 0));
#pragma unroll
            for (int jj = 0; jj < 32; jj += 2) {
              const bool anyc = (jj < 16) ? any0 : any1;
              const float mine_c0 = anyc ? __uint_as_float(v[jj]) : 0.0f, mine_c1 = anyc ? __uint_as_float(v[jj + 1]) : 0.0f;
              const float send = odd ? mine_c0 : mine_c1;                 // even lanes give away column c+1, odd lanes column c
              const float got = __shfl_xor_sync(0xffffffffu, send, 1);
              const __nv_bfloat162 pk = odd ? __floats2bfloat162_rn(got, mine_c1) : __floats2bfloat162_rn(mine_c0, got);
              if (valid && (jj < 16 || ncol == 32) && !(P.skip & 8)) dst32[((jj + (odd ? 1 : 0)) * M) >> 1] = *reinterpret_cast<const unsigned int*>(&pk);
            }
          } else if (valid) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              if (jj < 16 || ncol == 32) {
                float acc = ((jj < 16) ? any0 : any1) ? __uint_as_float(v[jj]) : 0.0f;
                if (!P.beta0) acc += __bfloat162float(dst[jj * M]);
                dst[jj * M] = __float2bfloat16_rn(acc);
              }
            }
          }
        }
      }
    }
    if (DBG && P.dbg != nullptr && bid == 0 && warp == 4 && lane == 0) { P.dbg[8] = w0; P.dbg[9] = XB_CLOCK() - tstart; }
  } else if (warp >= 12 && warp < 12 + kConvWarps) {
    // ========================================= converters =======================================
    // One work unit = 4 consecutive rows (m) of one k-pair: a 16-byte read of VNNI words, two 8-byte writes (rows k, k+1 of
    // the canonical operand). Lanes 0-15 of a half-warp take the 16 units of ONE (k-pair, 64-row atom): their reads are
    // two contiguous 128-byte runs and their writes fill one whole 128-byte row each -- no shared-memory bank conflicts
    // (the first version's 32-byte lane stride made every access 2-way conflicted: 18M of 30M wavefronts in ncu).
    const int cw = warp - 12, q = lane & 15, kp_lo = lane >> 4;
    int rs = 0, cs = 0; uint32_t rph = 0, cph = 0; long long w_r = 0, w_c = 0; const long long tstart = XB_CLOCK();
    for (long long i = 0; i < n_local; ++i) {
      for (int ks = 0; ks < NKS; ++ks) {
        XB_READY(w_r, raw_full + 8 * rs, rph); XB_TWAIT(w_r, mbar_wait_x(raw_full + 8 * rs, rph, P.spin & 4));
        XB_READY(w_c, can_empty + 8 * cs, cph ^ 1); XB_TWAIT(w_c, mbar_wait_x(can_empty + 8 * cs, cph ^ 1, P.spin & 4));
        const uint8_t* src = s_raw + (size_t)rs * A_STAGE;
        uint8_t* dst = s_can + (size_t)cs * A_STAGE;
        if (!(P.skip & 1)) {
#pragma unroll
          for (int it = 0; it < 32 / kConvWarps; ++it) {
            const int idx = cw * (32 / kConvWarps) + it;          // 0..31: (atom, pair of k-pairs)
            const int atom = idx & 1, kp = (idx >> 1) * 2 + kp_lo;
            const int mrow = atom * 64 + 4 * q, g = mrow / M, mm = mrow % M;
            const uint4 w = *reinterpret_cast<const uint4*>(src + ((size_t)(g * 32 + kp) * M + mm) * 4);
            uint2 ev, od;
            ev.x = __byte_perm(w.x, w.y, 0x5410); ev.y = __byte_perm(w.z, w.w, 0x5410);
            od.x = __byte_perm(w.x, w.y, 0x7632); od.y = __byte_perm(w.z, w.w, 0x7632);
            uint8_t* base = dst + (size_t)atom * (64 * 128);
            const int k0 = 2 * kp, k1 = 2 * kp + 1, ch = q >> 1, sub = (q & 1) << 3;
            *reinterpret_cast<uint2*>(base + k0 * 128 + ((ch ^ (k0 & 7)) << 4) + sub) = ev;
            *reinterpret_cast<uint2*>(base + k1 * 128 + ((ch ^ (k1 & 7)) << 4) + sub) = od;
          }
        }
        if (!(P.skip & 128)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) { mbar_arrive(can_full + 8 * cs); mbar_arrive(raw_empty + 8 * rs); }
        if (++rs == RS) { rs = 0; rph ^= 1; }
        if (++cs == CS) { cs = 0; cph ^= 1; }
      }
    }
    if (DBG && P.dbg != nullptr && bid == 0 && cw == 0 && lane == 0) { P.dbg[10] = w_r; P.dbg[11] = w_c; P.dbg[12] = XB_CLOCK() - tstart; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
  }
}


// ---- variant 2: two CTAs per SM -------------------------------------------------------------------------------------------
// The pipeline above is bound by hand-over LATENCY (a stage is occupied from the start of its conversion until its MMAs
// retire, ~2000+ cycles, while the HBM budget is ~670 cycles per k-step), not by any throughput: the cure that worked for the
// dense kernel is a second, independent pipeline on the same SM. To make two CTAs fit (<= ~112 KB of shared memory, 256 TMEM
// columns, <= 85 registers x 384 threads each) the VNNI -> canonical conversion is done IN PLACE (every converter thread pulls
// its 8 x 16 bytes into registers, the 128 converter threads meet at a named barrier, then write), which removes the second A
// ring, and each CTA keeps a single 256-column accumulator: while its epilogue drains, the sibling CTA computes.
//   warp 0 A producer | warp 2 B producer | warps 1,3 MMA issuers (column ownership as above) | warps 4-7 epilogue | warps 8-11 converters
constexpr int kThreads2 = 384;

template <int M>
__global__ void __launch_bounds__(kThreads2, 2)
bcsc_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const BcscTcParams P) {
  constexpr int G = 128 / M;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int AS = P.raw_stages, BS = P.b_stages;
  const int NP = P.nparts, NKS = P.nks, NL = P.nparts * P.nks;
  uint8_t* s_a = smem;
  uint8_t* s_b = s_a + (size_t)AS * A_STAGE;
  uint4* s_ops = (uint4*)(s_b + (size_t)BS * P.b_stage_bytes);
  unsigned int* s_lp = (unsigned int*)(s_ops + P.ops_cap + 1);
  unsigned int* s_wr = s_lp + NL + 1;
  unsigned char* s_any = (unsigned char*)(s_wr + 4 * NL);
  uint64_t* bars = (uint64_t*)(((uintptr_t)(s_any + P.nbc) + 15) & ~(uintptr_t)15);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t a_full = bar0, a_conv = a_full + 8 * 8, a_empty = a_conv + 8 * 8, b_full = a_empty + 8 * 8, b_empty = b_full + 8 * 8;
  const uint32_t t_full = b_empty + 8 * 8, t_empty = t_full + 8;
  uint32_t* tmem_word = (uint32_t*)(bars + 5 * 8 + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long Gd = gridDim.x, bid = blockIdx.x;
  const long long n_local = ((bid < P.ngroups) ? (P.ngroups - bid + Gd - 1) / Gd : 0) * NP;
  const uint32_t blk_bytes = (uint32_t)P.bn * P.bk * 2;
  {
    const unsigned int nnzb = P.list_ptr[NL];
    for (unsigned int i = threadIdx.x; i < nnzb && i < (unsigned int)P.ops_cap; i += blockDim.x) s_ops[i] = P.ops[i];
    for (int i = threadIdx.x; i < 4 * NL; i += blockDim.x) s_wr[i] = P.wranges[i];
    for (int i = threadIdx.x; i <= NL; i += blockDim.x) s_lp[i] = P.list_ptr[i];
    for (int i = threadIdx.x; i < P.nbc; i += blockDim.x) s_any[i] = (unsigned char)P.col_any[i];
  }
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_a) : "memory");
    for (int i = 0; i < AS; ++i) { mbar_init(a_full + 8 * i, 1); mbar_init(a_conv + 8 * i, 4); mbar_init(a_empty + 8 * i, 2); }
    for (int i = 0; i < BS; ++i) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 2); }
    mbar_init(t_full, 2); mbar_init(t_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_word)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_word;

  if (warp == 0) {
    if (lane == 0) {
      int as = 0; uint32_t aph = 0;
      for (long long i = 0; i < n_local; ++i) {
        const long long grp = bid + (i / NP) * Gd;
        for (int ks = 0; ks < NKS; ++ks) {
          mbar_wait(a_empty + 8 * as, aph ^ 1);
          mbar_expect_tx(a_full + 8 * as, (uint32_t)A_STAGE);
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                       :: "r"(smem_u32(s_a + (size_t)as * A_STAGE)), "l"(&map_a), "r"(0), "r"(ks * 32), "r"((int)(grp * G)), "r"(a_full + 8 * as) : "memory");
          if (++as == AS) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      int bs = 0; uint32_t bph = 0;
      for (long long i = 0; i < n_local; ++i) {
        const int l0 = (int)(i % NP) * NKS;
        for (int ks = 0; ks < NKS; ++ks) {
          const unsigned int e0 = s_lp[l0 + ks], e1 = s_lp[l0 + ks + 1];
          mbar_wait(b_empty + 8 * bs, bph ^ 1);
          mbar_expect_tx(b_full + 8 * bs, (e1 - e0) * blk_bytes);
          if (e1 > e0) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(s_b + (size_t)bs * P.b_stage_bytes)), "l"(P.b_packed + (size_t)e0 * blk_bytes), "r"((e1 - e0) * blk_bytes),
                            "r"(b_full + 8 * bs) : "memory");
          }
          if (++bs == BS) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const int mw = (warp == 1) ? 0 : 1;
    int as = 0, bs = 0; uint32_t aph = 0, bph = 0;
    const uint32_t a_hi = desc_hi(1024 >> 4, 2), b_hi = desc_hi(P.b_sbo16, P.b_layout);
    const uint32_t a_lo0 = desc_lo(smem_u32(s_a), (uint32_t)(64 * 128) >> 4);
    const uint32_t b_lo0 = desc_lo(smem_u32(s_b), 1);
    const int ksteps = P.ksteps;
    for (long long i = 0; i < n_local; ++i) {
      mbar_wait(t_empty, (uint32_t)((i & 1) ^ 1));
      tc_fence_after();
      const int l0 = (int)(i % NP) * NKS;
      for (int ks = 0; ks < NKS; ++ks) {
        const unsigned int rng = s_wr[4 * (l0 + ks) + mw], ob = rng & 0xFFFFu, on = rng >> 16;
        uint4 op = s_ops[ob];
        mbar_wait(a_conv + 8 * as, aph);
        mbar_wait(b_full + 8 * bs, bph);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + (uint32_t)as * (A_STAGE >> 4);
        const uint32_t b_stage_lo = b_lo0 + (uint32_t)bs * ((uint32_t)P.b_stage_bytes >> 4);
        for (unsigned int o = 0; o < on; ++o) {
          const uint4 nxt = s_ops[ob + o + 1];
          const uint32_t d = tmem_base + (op.x & 0xFFFFu), a_op = a_lo + (op.x >> 16), b_lo = b_stage_lo + op.y, idesc = P.idesc | op.z;
          uint32_t accumulate = op.w;
          for (int kk = 0; kk < ksteps; ++kk) {
            if (leader) umma_f16(d, desc64(a_hi, a_op + kk * (2048 >> 4)), desc64(b_hi, b_lo + kk * (32 >> 4)), idesc, accumulate);
            accumulate = 1;
          }
          op = nxt;
        }
        if (leader) { umma_commit(a_empty + 8 * as); umma_commit(b_empty + 8 * bs); }
        __syncwarp();
        if (++as == AS) { as = 0; aph ^= 1; }
        if (++bs == BS) { bs = 0; bph ^= 1; }
      }
      if (leader) umma_commit(t_full);
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 8) {
    const int q = warp & 3;
    const int row = 32 * q + lane;
    const int mbl = row / M, m = row % M;
    const bool bn32 = (P.bn % 32) == 0;
    for (long long i = 0; i < n_local; ++i) {
      const long long grp = bid + (i / NP) * Gd;
      const long long mb = grp * G + mbl;
      const bool valid = mb < P.m_blocks;
      const int part = (int)(i % NP), pc0 = part * P.part_cols;
      const int pcols = (P.ncols - pc0 < P.part_cols) ? (P.ncols - pc0) : P.part_cols;
      const int nchunks = (pcols + 31) / 32;
      __nv_bfloat16* cblk = reinterpret_cast<__nv_bfloat16*>(P.c) + ((size_t)mb * P.ncols + pc0) * M + m;
      mbar_wait(t_full, (uint32_t)(i & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int ch = 0; ch < nchunks; ++ch) {
        const int c0 = ch * 32;
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        if (ch + 1 == nchunks) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(t_empty); }
        __nv_bfloat16* dst = cblk + (size_t)c0 * M;
        const int ncol = (pcols - c0 < 32) ? (pcols - c0) : 32;
        const int jb = (pc0 + c0) / P.bn;
        const bool any0 = s_any[jb] != 0, any1 = bn32 ? any0 : (s_any[(pc0 + c0 + 16) / P.bn < P.nbc ? (pc0 + c0 + 16) / P.bn : jb] != 0);
        if (P.beta0 && (M % 2) == 0) {
          const bool odd = lane & 1;
          unsigned int* dst32 = reinterpret_cast<unsigned int*>(dst - (odd ? 1 : 0));
#pragma unroll
          for (int jj = 0; jj < 32; jj += 2) {
            const bool anyc = (jj < 16) ? any0 : any1;
            const float mine_c0 = anyc ? __uint_as_float(v[jj]) : 0.0f, mine_c1 = anyc ? __uint_as_float(v[jj + 1]) : 0.0f;
            const float send = odd ? mine_c0 : mine_c1;
            const float got = __shfl_xor_sync(0xffffffffu, send, 1);
            const __nv_bfloat162 pk = odd ? __floats2bfloat162_rn(got, mine_c1) : __floats2bfloat162_rn(mine_c0, got);
            if (valid && (jj < 16 || ncol == 32)) dst32[((jj + (odd ? 1 : 0)) * M) >> 1] = *reinterpret_cast<const unsigned int*>(&pk);
          }
        } else if (valid) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            if (jj < 16 || ncol == 32) {
              float acc = ((jj < 16) ? any0 : any1) ? __uint_as_float(v[jj]) : 0.0f;
              if (!P.beta0) acc += __bfloat162float(dst[jj * M]);
              dst[jj * M] = __float2bfloat16_rn(acc);
            }
          }
        }
      }
      if (nchunks == 0) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(t_empty); }
    }
  } else if (warp >= 8 && warp < 12) {
    // in-place conversion: 1024 units (4 rows x 1 k-pair) per stage, 8 per thread; same lane mapping as above (conflict-free)
    const int cw = warp - 8, q = lane & 15, kp_lo = lane >> 4;
    int as = 0; uint32_t aph = 0;
    for (long long i = 0; i < n_local; ++i) {
      for (int ks = 0; ks < NKS; ++ks) {
        mbar_wait(a_full + 8 * as, aph);
        uint8_t* buf = s_a + (size_t)as * A_STAGE;
        uint4 w[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int idx = cw * 8 + it, atom = idx & 1, kp = (idx >> 1) * 2 + kp_lo;
          const int mrow = atom * 64 + 4 * q, g = mrow / M, mm = mrow % M;
          w[it] = *reinterpret_cast<const uint4*>(buf + ((size_t)(g * 32 + kp) * M + mm) * 4);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");          // every converter thread holds its share of the raw stage
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int idx = cw * 8 + it, atom = idx & 1, kp = (idx >> 1) * 2 + kp_lo;
          uint2 ev, od;
          ev.x = __byte_perm(w[it].x, w[it].y, 0x5410); ev.y = __byte_perm(w[it].z, w[it].w, 0x5410);
          od.x = __byte_perm(w[it].x, w[it].y, 0x7632); od.y = __byte_perm(w[it].z, w[it].w, 0x7632);
          uint8_t* base = buf + (size_t)atom * (64 * 128);
          const int k0 = 2 * kp, k1 = 2 * kp + 1, ch = q >> 1, sub = (q & 1) << 3;
          *reinterpret_cast<uint2*>(base + k0 * 128 + ((ch ^ (k0 & 7)) << 4) + sub) = ev;
          *reinterpret_cast<uint2*>(base + k1 * 128 + ((ch ^ (k1 & 7)) << 4) + sub) = od;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(a_conv + 8 * as);
        if (++as == AS) { as = 0; aph ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(256u) : "memory");
  }
}

template <int M>
cudaError_t launch_two(long long grid, size_t smem, cudaStream_t stream, const CUtensorMap& ma, const BcscTcParams& P) {
  static int attr_set = 0;
  if (!attr_set) { cudaFuncSetAttribute(bcsc_tc2_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024); attr_set = 1; }
  bcsc_tc2_kernel<M><<<(unsigned int)grid, kThreads2, smem, stream>>>(ma, P);
  return cudaGetLastError();
}

// ---- variant 3 (opt-in, LIBXSMM_B200_BCSC_V3=1): the A operand lives in TENSOR MEMORY ---------------------------------------
// For kind::f16 a 32-bit TMEM cell of the A operand holds two consecutive k of one row -- which is exactly one VNNI2 word.
// So the "conversion" degenerates to a transposing copy: thread = row m reads its 32 words of a k-step from the raw TMA
// stage (conflict-free) and writes them with ONE tcgen05.st.32x32b.x32 into an A ring in TMEM; the MMAs then take A from
// TMEM (no 4 KB shared-memory read of A per instruction, no canonical ring, no PRMT, no proxy fence). TMEM plan: two
// accumulator slots of 192 columns (column parts of <= 192) + four A stages of 32 columns = 512.
constexpr int kA3Stages = 4, kD3Cols = 192;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
               "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
               "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                  "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                  "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                  "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

template <int M>
__global__ void __launch_bounds__(kThreads, 1)
bcsc_tc3_kernel(const __grid_constant__ CUtensorMap map_a, const BcscTcParams P) {
  constexpr int G = 128 / M;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int RS = P.raw_stages, BS = P.b_stages, TS = kA3Stages;
  const int NP = P.nparts, NKS = P.nks, NL = P.nparts * P.nks;
  uint8_t* s_raw = smem;
  uint8_t* s_b = s_raw + (size_t)RS * A_STAGE;
  uint4* s_ops = (uint4*)(s_b + (size_t)BS * P.b_stage_bytes);
  unsigned int* s_lp = (unsigned int*)(s_ops + P.ops_cap + 1);
  unsigned int* s_wr = s_lp + NL + 1;
  unsigned char* s_any = (unsigned char*)(s_wr + 4 * NL);
  uint64_t* bars = (uint64_t*)(((uintptr_t)(s_any + P.nbc) + 15) & ~(uintptr_t)15);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t raw_full = bar0, raw_empty = raw_full + 8 * 16, ta_full = raw_empty + 8 * 16, ta_empty = ta_full + 8 * 8;
  const uint32_t b_full = ta_empty + 8 * 8, b_empty = b_full + 8 * 16, t_full = b_empty + 8 * 16, t_empty = t_full + 8 * 2;
  uint32_t* tmem_word = (uint32_t*)(bars + 4 * 16 + 2 * 8 + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long Gd = gridDim.x, bid = blockIdx.x;
  const long long n_local = ((bid < P.ngroups) ? (P.ngroups - bid + Gd - 1) / Gd : 0) * NP;
  const uint32_t blk_bytes = (uint32_t)P.bn * P.bk * 2;
  {
    const unsigned int nnzb = P.list_ptr[NL];
    for (unsigned int i = threadIdx.x; i < nnzb && i < (unsigned int)P.ops_cap; i += blockDim.x) s_ops[i] = P.ops[i];
    for (int i = threadIdx.x; i < 4 * NL; i += blockDim.x) s_wr[i] = P.wranges[i];
    for (int i = threadIdx.x; i <= NL; i += blockDim.x) s_lp[i] = P.list_ptr[i];
    for (int i = threadIdx.x; i < P.nbc; i += blockDim.x) s_any[i] = (unsigned char)P.col_any[i];
  }
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_a) : "memory");
    for (int i = 0; i < RS; ++i) { mbar_init(raw_full + 8 * i, 1); mbar_init(raw_empty + 8 * i, 4); }
    for (int i = 0; i < TS; ++i) { mbar_init(ta_full + 8 * i, 4); mbar_init(ta_empty + 8 * i, (uint32_t)P.mma_warps); }
    for (int i = 0; i < BS; ++i) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, (uint32_t)P.mma_warps); }
    for (int i = 0; i < 2; ++i) { mbar_init(t_full + 8 * i, (uint32_t)P.mma_warps); mbar_init(t_empty + 8 * i, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_word)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_word;
  const uint32_t tmem_a = tmem_base + 2u * kD3Cols;          // A ring: columns 384..511

  if (warp == 0) {
    if (lane == 0) {
      int rs = 0; uint32_t rph = 0;
      for (long long i = 0; i < n_local; ++i) {
        const long long grp = bid + (i / NP) * Gd;
        for (int ks = 0; ks < NKS; ++ks) {
          mbar_wait(raw_empty + 8 * rs, rph ^ 1);
          mbar_expect_tx(raw_full + 8 * rs, (uint32_t)A_STAGE);
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                       :: "r"(smem_u32(s_raw + (size_t)rs * A_STAGE)), "l"(&map_a), "r"(0), "r"(ks * 32), "r"((int)(grp * G)), "r"(raw_full + 8 * rs) : "memory");
          if (++rs == RS) { rs = 0; rph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      int bs = 0; uint32_t bph = 0;
      for (long long i = 0; i < n_local; ++i) {
        const int l0 = (int)(i % NP) * NKS;
        for (int ks = 0; ks < NKS; ++ks) {
          const unsigned int e0 = s_lp[l0 + ks], e1 = s_lp[l0 + ks + 1];
          mbar_wait(b_empty + 8 * bs, bph ^ 1);
          mbar_expect_tx(b_full + 8 * bs, (e1 - e0) * blk_bytes);
          if (e1 > e0) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(s_b + (size_t)bs * P.b_stage_bytes)), "l"(P.b_packed + (size_t)e0 * blk_bytes), "r"((e1 - e0) * blk_bytes),
                            "r"(b_full + 8 * bs) : "memory");
          }
          if (++bs == BS) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp == 1 || (warp == 3 && P.mma_warps >= 2) || (warp >= 20 && warp - 18 < P.mma_warps)) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const int mw = (warp == 1) ? 0 : ((warp == 3) ? 1 : warp - 18);
    int ts = 0, bs = 0; uint32_t tph = 0, bph = 0;
    const uint32_t b_hi = desc_hi(P.b_sbo16, P.b_layout);
    const uint32_t b_lo0 = desc_lo(smem_u32(s_b), 1);
    const uint32_t idesc0 = P.idesc & ~(1u << 15);           // A from TMEM: rows in lanes, k along the columns (K-major)
    const int ksteps = P.ksteps;
    for (long long i = 0; i < n_local; ++i) {
      const int slot = (int)(i & 1);
      mbar_wait(t_empty + 8 * slot, (uint32_t)(((i >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint32_t d_base = tmem_base + (uint32_t)(slot * kD3Cols);
      const int l0 = (int)(i % NP) * NKS;
      for (int ks = 0; ks < NKS; ++ks) {
        const unsigned int rng = s_wr[4 * (l0 + ks) + mw], ob = rng & 0xFFFFu, on = rng >> 16;
        uint4 op = s_ops[ob];
        mbar_wait(ta_full + 8 * ts, tph);
        mbar_wait(b_full + 8 * bs, bph);
        tc_fence_after();
        const uint32_t a_stage = tmem_a + (uint32_t)ts * 32u;
        const uint32_t b_stage_lo = b_lo0 + (uint32_t)bs * ((uint32_t)P.b_stage_bytes >> 4);
        for (unsigned int o = 0; o < on; ++o) {
          const uint4 nxt = s_ops[ob + o + 1];
          const uint32_t d = d_base + (op.x & 0xFFFFu), a_col = a_stage + ((op.x >> 16) >> 4), b_lo = b_stage_lo + op.y, idesc = idesc0 | op.z;
          uint32_t accumulate = op.w;
          for (int kk = 0; kk < ksteps; ++kk) {
            if (leader) umma_f16_ts(d, a_col + (uint32_t)kk * 8u, desc64(b_hi, b_lo + kk * (32 >> 4)), idesc, accumulate);
            accumulate = 1;
          }
          op = nxt;
        }
        if (leader) { umma_commit(ta_empty + 8 * ts); umma_commit(b_empty + 8 * bs); }
        __syncwarp();
        if (++ts == TS) { ts = 0; tph ^= 1; }
        if (++bs == BS) { bs = 0; bph ^= 1; }
      }
      if (leader) umma_commit(t_full + 8 * slot);
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 12) {
    const int q = warp & 3, half = (warp - 4) >> 2;
    const int row = 32 * q + lane;
    const int mbl = row / M, m = row % M;
    const bool bn32 = (P.bn % 32) == 0;
    for (long long i = 0; i < n_local; ++i) {
      const int slot = (int)(i & 1);
      const long long grp = bid + (i / NP) * Gd;
      const long long mb = grp * G + mbl;
      const bool valid = mb < P.m_blocks;
      const int part = (int)(i % NP), pc0 = part * P.part_cols;
      const int pcols = (P.ncols - pc0 < P.part_cols) ? (P.ncols - pc0) : P.part_cols;
      const int nchunks = (pcols + 31) / 32, cbeg = half ? (nchunks + 1) / 2 : 0, cend = half ? nchunks : (nchunks + 1) / 2;
      __nv_bfloat16* cblk = reinterpret_cast<__nv_bfloat16*>(P.c) + ((size_t)mb * P.ncols + pc0) * M + m;
      mbar_wait(t_full + 8 * slot, (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(slot * kD3Cols) + ((uint32_t)(q * 32) << 16);
      if (cbeg == cend) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(t_empty + 8 * slot); }
      for (int ch = cbeg; ch < cend; ++ch) {
        const int c0 = ch * 32;
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        if (ch + 1 == cend) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(t_empty + 8 * slot); }
        __nv_bfloat16* dst = cblk + (size_t)c0 * M;
        const int ncol = (pcols - c0 < 32) ? (pcols - c0) : 32;
        const int jb = (pc0 + c0) / P.bn;
        const bool any0 = s_any[jb] != 0, any1 = bn32 ? any0 : (s_any[(pc0 + c0 + 16) / P.bn < P.nbc ? (pc0 + c0 + 16) / P.bn : jb] != 0);
        if (P.beta0 && (M % 2) == 0) {
          const bool odd = lane & 1;
          unsigned int* dst32 = reinterpret_cast<unsigned int*>(dst - (odd ? 1 : 0));
#pragma unroll
          for (int jj = 0; jj < 32; jj += 2) {
            const bool anyc = (jj < 16) ? any0 : any1;
            const float mine_c0 = anyc ? __uint_as_float(v[jj]) : 0.0f, mine_c1 = anyc ? __uint_as_float(v[jj + 1]) : 0.0f;
            const float send = odd ? mine_c0 : mine_c1;
            const float got = __shfl_xor_sync(0xffffffffu, send, 1);
            const __nv_bfloat162 pk = odd ? __floats2bfloat162_rn(got, mine_c1) : __floats2bfloat162_rn(mine_c0, got);
            if (valid && (jj < 16 || ncol == 32)) dst32[((jj + (odd ? 1 : 0)) * M) >> 1] = *reinterpret_cast<const unsigned int*>(&pk);
          }
        } else if (valid) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            if (jj < 16 || ncol == 32) {
              float acc = ((jj < 16) ? any0 : any1) ? __uint_as_float(v[jj]) : 0.0f;
              if (!P.beta0) acc += __bfloat162float(dst[jj * M]);
              dst[jj * M] = __float2bfloat16_rn(acc);
            }
          }
        }
      }
    }
  } else if (warp >= 12 && warp < 16) {
    // transposing copy raw VNNI words -> TMEM A stage: this warp owns rows 32q .. 32q+31 (its TMEM lane quadrant)
    const int q = warp & 3, mrow = 32 * q + lane, g = mrow / M, mm = mrow % M;
    int rs = 0, ts = 0; uint32_t rph = 0, tph = 0;
    for (long long i = 0; i < n_local; ++i) {
      for (int ks = 0; ks < NKS; ++ks) {
        mbar_wait(raw_full + 8 * rs, rph);
        mbar_wait(ta_empty + 8 * ts, tph ^ 1);
        tc_fence_after();
        const unsigned int* src = reinterpret_cast<const unsigned int*>(s_raw + (size_t)rs * A_STAGE) + (size_t)g * 32 * M + mm;
        uint32_t w[32];
#pragma unroll
        for (int kp = 0; kp < 32; ++kp) w[kp] = src[(size_t)kp * M];
        tmem_st32(tmem_a + (uint32_t)ts * 32u + ((uint32_t)(q * 32) << 16), w);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(ta_full + 8 * ts); mbar_arrive(raw_empty + 8 * rs); }
        if (++rs == RS) { rs = 0; rph ^= 1; }
        if (++ts == TS) { ts = 0; tph ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int M>
cudaError_t launch_three(long long grid, size_t smem, cudaStream_t stream, const CUtensorMap& ma, const BcscTcParams& P) {
  static int attr_set = 0;
  if (!attr_set) { cudaFuncSetAttribute(bcsc_tc3_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_set = 1; }
  bcsc_tc3_kernel<M><<<(unsigned int)grid, kThreads, smem, stream>>>(ma, P);
  return cudaGetLastError();
}

int g_sms = 0;

template <int M, bool DBG>
cudaError_t launch_one(long long grid, size_t smem, cudaStream_t stream, const CUtensorMap& ma, const BcscTcParams& P) {
  static int attr_set = 0;
  if (!attr_set) { cudaFuncSetAttribute(bcsc_tc_kernel<M, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_set = 1; }
  bcsc_tc_kernel<M, DBG><<<(unsigned int)grid, kThreads, smem, stream>>>(ma, P);
  return cudaGetLastError();
}

int env_int(const char* name, int lo, int hi, int dflt) {
  const char* e = getenv(name);
  if (e == nullptr || *e == 0) return dflt;
  const int v = atoi(e);
  return (v < lo || v > hi) ? dflt : v;
}

}  // namespace

// returns 0 if launched, <0 if this descriptor/problem is not served by the tensor-core kernel (caller falls back)
extern "C" int xb_bcsc_tc_launch(xb_sparse_desc* d, const void* a, const void* b_vals, const unsigned int* colptr,
                                 const unsigned int* rowidx, unsigned long long n_blocks, unsigned int nnzb, void* c)
{
  const unsigned int bad = LIBXSMM_GEMM_FLAG_TRANS_A | LIBXSMM_GEMM_FLAG_TRANS_B | LIBXSMM_GEMM_FLAG_VNNI_B;
  const int M = d->packed_width, K = d->k, bk = d->bk, bn = d->bn;
  if (getenv("LIBXSMM_B200_BCSC_SIMT") != nullptr) return -1;
  if (d->ta != LIBXSMM_DATATYPE_BF16 || d->tb != LIBXSMM_DATATYPE_BF16 || d->tc != LIBXSMM_DATATYPE_BF16) return -1;
  if ((d->flags & bad) != 0 || (d->flags & LIBXSMM_GEMM_FLAG_VNNI_A) == 0) return -1;
  if (!(M == 16 || M == 32 || M == 64 || M == 128)) return -1;
  if (!(bk == 16 || bk == 32 || bk == 64) || (bn % 16) != 0 || bn > 256 || bn < 16 || K < 64 || (K % bk) != 0) return -1;
  const long long ncols = (long long)n_blocks * bn;
  const int nkb = K / bk, nks = (K + 63) / 64;
  // nnzb == 0: unknown on the host (pattern lives in device memory): bound it by the dense block count; the kernels
  // read the true count from the prep output
  if (nnzb == 0) nnzb = (unsigned int)((unsigned long long)n_blocks * nkb < 2047ull ? n_blocks * nkb : 2047ull);
  if (ncols > 512 || nkb > 255 || nnzb > 2047 || n_blocks > 255 || (unsigned long long)n_blocks * nkb > 2047ull) return -1;
  // column parts: an accumulator of <= 256 columns leaves room for a second one in TMEM, so the epilogue of one work item
  // overlaps the MMAs of the next (a single 512-column accumulator serialises the two phases)
  int bpp = (int)n_blocks, nparts = 1;
  const bool v3 = env_int("LIBXSMM_B200_BCSC_V3", 0, 1, 0) == 1 && bn <= kD3Cols;      // opt-in: A operand in tensor memory
  if (v3) { bpp = kD3Cols / bn; if (bpp > (int)n_blocks) bpp = (int)n_blocks; nparts = ((int)n_blocks + bpp - 1) / bpp; }
  else if (ncols > 256) { bpp = 256 / bn; nparts = ((int)n_blocks + bpp - 1) / bpp; }
  if (env_int("LIBXSMM_B200_BCSC_PARTS", 1, 1, 0) == 1) { bpp = (int)n_blocks; nparts = 1; }
  const int nl = nparts * nks;
  if (nl > kMaxLists - 1) return -1;
  if ((((uintptr_t)a | (uintptr_t)b_vals | (uintptr_t)c) & 15) != 0) return -1;
  xb_encode_tiled_fn enc = xb_tma_encoder();
  if (enc == nullptr) return -1;
  if (g_sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev); if (g_sms <= 0) g_sms = 148; }
  cudaStream_t stream = (cudaStream_t)xb_rt_stream();

  BcscTcParams P; memset(&P, 0, sizeof(P));
  const int G = 128 / M, blk_bytes = bn * bk * 2;
  P.bn = bn; P.bk = bk; P.nbc = (int)n_blocks; P.nks = nks; P.ksteps = bk / 16;
  P.m_blocks = d->m; P.ngroups = (d->m + G - 1) / G;
  P.ncols = (int)ncols; P.nparts = nparts; P.part_cols = (nparts == 1) ? (int)ncols : bpp * bn;
  P.slot_cols = (P.part_cols + 31) & ~31; P.nslot = 512 / P.slot_cols; if (P.nslot > 4) P.nslot = 4;
  // shared-memory plan: a B stage holds one whole k-step of a part (worst case: every block present)
  P.b_stage_bytes = bpp * (64 / bk) * blk_bytes;
  const size_t meta = (size_t)(kMaxEntries + 1) * 16 + (5 * kMaxLists + 8) * 4 + 1024 + (6 * kMaxStages + 2 * kMaxSlots) * 8 + 64 + 1024;
  const size_t budget = 224 * 1024 - meta;
  P.b_stages = env_int("LIBXSMM_B200_BCSC_BST", 2, 16, 3); P.can_stages = env_int("LIBXSMM_B200_BCSC_CAN", 2, 8, 3);
  while (P.b_stages > 2 && (size_t)P.b_stages * P.b_stage_bytes + (size_t)(P.can_stages + 2) * A_STAGE > budget) --P.b_stages;
  if ((size_t)P.b_stages * P.b_stage_bytes + (size_t)(P.can_stages + 2) * A_STAGE > budget) return -1;
  P.raw_stages = (int)((budget - (size_t)P.b_stages * P.b_stage_bytes - (size_t)P.can_stages * A_STAGE) / A_STAGE);
  if (P.raw_stages > 8) P.raw_stages = 8;
  P.raw_stages = env_int("LIBXSMM_B200_BCSC_RAW", 2, P.raw_stages, P.raw_stages);
  P.mma_warps = env_int("LIBXSMM_B200_BCSC_MMAW", 1, 4, 4);
  // variant 2 (two CTAs per SM): one 256-column accumulator, two stages per ring, everything within 112 KB
  P.ops_cap = (int)nnzb;
  const size_t meta2 = ((size_t)P.ops_cap + 1) * 16 + ((size_t)5 * nl + 8) * 4 + (size_t)n_blocks + 16 + (5 * 8 + 2) * 8 + 64 + 1024;
  const int a2 = env_int("LIBXSMM_B200_BCSC_V2_AST", 2, 4, 2), b2 = env_int("LIBXSMM_B200_BCSC_V2_BST", 2, 4, 2);
  const size_t smem2 = meta2 + (size_t)a2 * A_STAGE + (size_t)b2 * P.b_stage_bytes;
  // measured (BASELINE size): variant 2 0.226 ms, variant 1 0.221 ms -- co-residency does not help here, so it stays opt-in
  const bool v2 = !v3 && env_int("LIBXSMM_B200_BCSC_V2", 0, 1, 0) == 1 && P.slot_cols <= 256 && smem2 <= 112 * 1024;
  const size_t meta3 = ((size_t)P.ops_cap + 1) * 16 + ((size_t)5 * nl + 8) * 4 + (size_t)n_blocks + 16 + (4 * 16 + 2 * 8 + 4) * 8 + 64 + 1024;
  size_t smem3 = 0;
  if (v3) {
    P.mma_warps = env_int("LIBXSMM_B200_BCSC_MMAW", 1, 4, 4);
    P.raw_stages = env_int("LIBXSMM_B200_BCSC_V3_RAW", 2, 8, 4); P.b_stages = env_int("LIBXSMM_B200_BCSC_V3_BST", 2, 8, 5);
    while (P.b_stages > 2 && meta3 + (size_t)P.raw_stages * A_STAGE + (size_t)P.b_stages * P.b_stage_bytes > 224 * 1024) --P.b_stages;
    smem3 = meta3 + (size_t)P.raw_stages * A_STAGE + (size_t)P.b_stages * P.b_stage_bytes;
    if (smem3 > 224 * 1024) return -1;
  }
  if (v2) { P.mma_warps = 2; P.raw_stages = a2; P.b_stages = b2; }
  const int cpw = (bpp + P.mma_warps - 1) / P.mma_warps;

  // per-handle device buffer: cached pattern + visiting order (the handle is caller-owned; reused across calls)
  const BcscIdxLayout L((unsigned int)n_blocks, (unsigned int)nl, nnzb);
  const size_t need = (size_t)L.total * sizeof(unsigned int);
  if (d->d_idx == nullptr || d->nrows < need) {
    if (d->d_idx != nullptr) { cudaStreamSynchronize(stream); cudaFree(d->d_idx); }
    if (cudaMalloc((void**)&d->d_idx, need) != cudaSuccess) { d->d_idx = nullptr; d->nrows = 0; (void)cudaGetLastError(); return -1; }
    cudaMemsetAsync(d->d_idx, 0, 64, stream);       // header: no cached pattern yet
    d->nrows = (unsigned int)need;
  }
  unsigned int* buf = d->d_idx;
  bcsc_prep_kernel<<<1, 256, 0, stream>>>(colptr, rowidx, (int)n_blocks, nkb, bn, bk, nparts, bpp, cpw, nnzb, buf);
  xb_rt_count_launch();
  // per-handle buffer for the repacked B values (capacity in d->nnz, bytes)
  const size_t bneed = (size_t)nnzb * blk_bytes;
  if (d->d_val == nullptr || d->nnz < bneed) {
    if (d->d_val != nullptr) { cudaStreamSynchronize(stream); cudaFree(d->d_val); }
    if (cudaMalloc(&d->d_val, bneed) != cudaSuccess) { d->d_val = nullptr; d->nnz = 0; (void)cudaGetLastError(); return -1; }
    d->nnz = (unsigned int)bneed;
  }
  {
    const unsigned long long chunks = (unsigned long long)bneed / 16;
    const unsigned int pgrid = (unsigned int)((chunks + 255) / 256 < 1024 ? (chunks + 255) / 256 : 1024);
    bcsc_pack_b_kernel<<<pgrid ? pgrid : 1, 256, 0, stream>>>((const uint4*)b_vals, buf + L.entries, buf + L.list_ptr, nl, (uint4*)d->d_val, bn, bk);
    xb_rt_count_launch();
  }
  P.b_packed = (const char*)d->d_val;
  P.list_ptr = buf + L.list_ptr; P.wranges = buf + L.wranges; P.ops = (const uint4*)(buf + L.ops); P.col_any = buf + L.col_any;
  P.c = (char*)c; P.beta0 = d->beta0;
  P.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(128 >> 4) << 24);   // N is filled in per (merged) operation
  P.b_layout = (bk == 16) ? 6u : ((bk == 32) ? 4u : 2u);
  P.b_sbo16 = (uint32_t)(8 * bk * 2) >> 4;
  P.sleep = env_int("LIBXSMM_B200_BCSC_SLEEP", 0, 1, 1); P.spin = env_int("LIBXSMM_B200_BCSC_SPIN", 0, 15, 0);
  P.skip = env_int("LIBXSMM_B200_BCSC_SKIP", 0, 255, 0);
  P.b_cpasync = env_int("LIBXSMM_B200_BCSC_BCPASYNC", 0, 1, 0);
  P.dbg = nullptr;
  { const char* e = getenv("LIBXSMM_B200_BCSC_DEBUG"); if (e && *e) P.dbg = (long long*)(uintptr_t)strtoull(e, nullptr, 0); }

  CUtensorMap map_a;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)M, (cuuint64_t)(K / 2), (cuuint64_t)d->m};
    const cuuint64_t strides[2] = {(cuuint64_t)M * 4, (cuuint64_t)K * M * 2};
    const cuuint32_t box[3] = {(cuuint32_t)M, 32u, (cuuint32_t)G};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (CUDA_SUCCESS != enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void*)a, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return -1;
  }
  const size_t smem = v2 ? smem2 : meta + (size_t)(P.raw_stages + P.can_stages) * A_STAGE + (size_t)P.b_stages * P.b_stage_bytes;
  const long long grid = v2 ? (P.ngroups < 2ll * g_sms ? P.ngroups : 2ll * g_sms) : (P.ngroups < g_sms ? P.ngroups : g_sms);
  cudaError_t e = cudaErrorInvalidValue;
  if (v3 && M == 16) e = launch_three<16>(grid, smem3, stream, map_a, P);
  else if (v3 && M == 32) e = launch_three<32>(grid, smem3, stream, map_a, P);
  else if (v3 && M == 64) e = launch_three<64>(grid, smem3, stream, map_a, P);
  else if (v3 && M == 128) e = launch_three<128>(grid, smem3, stream, map_a, P);
  else if (v2 && M == 16) e = launch_two<16>(grid, smem, stream, map_a, P);
  else if (v2 && M == 32) e = launch_two<32>(grid, smem, stream, map_a, P);
  else if (v2 && M == 64) e = launch_two<64>(grid, smem, stream, map_a, P);
  else if (v2 && M == 128) e = launch_two<128>(grid, smem, stream, map_a, P);
  else if (P.dbg != nullptr && M == 32) e = launch_one<32, true>(grid, smem, stream, map_a, P);
  else if (M == 16) e = launch_one<16, false>(grid, smem, stream, map_a, P);
  else if (M == 32) e = launch_one<32, false>(grid, smem, stream, map_a, P);
  else if (M == 64) e = launch_one<64, false>(grid, smem, stream, map_a, P);
  else if (M == 128) e = launch_one<128, false>(grid, smem, stream, map_a, P);
  xb_rt_count_launch();
  if (e != cudaSuccess) { xb_rt_note_error((int)e, "bcsc_tc"); return (int)e; }
  return 0;
}
