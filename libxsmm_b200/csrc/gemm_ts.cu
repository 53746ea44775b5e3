// libxsmm_b200 -- batched small-tile GEMM/BRGEMM with a VNNI-packed A operand on the tcgen05 tensor cores (sm_100a):
// the reference's canonical low-precision layouts (src/generator_gemm_reference_impl.c:1452-1555 int8 VNNI4,
// :2127-2170 / :2367-2419 bf16 VNNI2, :2025-2126 f16) -- A[(k/v)*lda*v + m*v + k%v], v = 4 (8-bit) or 2 (16-bit).
//
//   C_t(m x n) = beta * C_t + sum_{r < br} A_{t,r}(m x k) * B_{t,r}(k x n)      for t < count tiles
//
// Why a second tensor-core kernel. gemm_tc.cu takes A straight from the caller's column-major buffer as an "MN-major"
// shared-memory operand. A VNNI-packed A has v consecutive k of one row in one 32-bit word -- no shared-memory operand
// layout describes that. But it is EXACTLY the layout of an A operand held in TENSOR MEMORY: for kind::f16 a 32-bit TMEM
// cell holds two consecutive k of one row, for kind::i8 four. So A needs no conversion, only a transposing copy
// (thread = row m reads its words of a k-chunk from the raw TMA stage and writes them with one tcgen05.st), and the MMA
// runs in TS form (A from TMEM, B from shared memory). B[n*ldb + k] is K-contiguous: the "K-major" operand, SWIZZLE_128B
// directly from TMA. Integer sums are exact in any order, so the int8 path stays BIT-IDENTICAL to the reference.
//
// Per CTA (persistent, tiles round-robin; several CTAs per SM share the 512 TMEM columns):
//   warp 0     TMA producer: per (tile, r, k-chunk) one box of raw A words (m words x 32 word-rows) and one of B
//              (128 bytes of k x n rows), ring of S stages
//   warp 1     MMA issuer  : tcgen05.mma.cta_group::1.kind::{f16,i8} [D], [A in TMEM], B descriptor; M = 128 always
//              (rows >= m carry whatever the stage holds and are never stored)
//   warps 2-5  epilogue    : tcgen05.ld 32x32b -> (+ beta*C) -> convert -> coalesced column stores (lane = row)
//   warps 6-9  copy        : raw stage -> TMEM A slot of the stage (one warp per TMEM lane quadrant)
// TMEM plan: NS accumulator slots of slot_cols columns, then S A-slots of 32 columns (A slot = ring stage).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "xb_internal.h"
#include "xb_device.cuh"
#include "xb_tma.cuh"
#include "xb_epilogue.cuh"

namespace {

struct TsParams {
  int m, n, k, np;              // np: n rounded up to 16
  int kchunks, kc_elems;        // k elements per stage: 128 bytes of a B row (64 for 16-bit, 128 for 8-bit)
  int kinst;                    // k per MMA instruction: 16 (16-bit) or 32 (8-bit)
  int stages, a_bytes, stage_bytes, tx_bytes;   // tx_bytes: what the two boxes of a stage deliver (the A region is padded to 1024)
  int nslot, slot_cols, tmem_cols, a_col0;
  unsigned long long br;
  long long count;
  char* c; long long tile_stride_c, ldc;
  int c_type, a_type, beta0, is_i8;
  int ep_mode, c_esz;           // xb_epilogue.cuh
  int grp, mp;                  // small tiles: grp tiles share one instruction (tile g*grp + j owns rows j*mp.. and columns j*np..); grp = 1: mp = 128
  long long groups;             // ceil(count / grp): the unit the CTAs walk
  float scf;
  uint32_t idesc;
  int skip;                     // diagnostic (LIBXSMM_B200_TS_SKIP): 1 no C stores, 2 no MMAs, 4 no TMEM copy, 8 no TMA loads
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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
               :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
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
constexpr int kTsgThreads = 320;     // 10 warps
constexpr int MAX_S = 8, MAX_NS = 4;

__global__ void __launch_bounds__(kTsgThreads, 3)      // <= 68 registers: three CTAs per SM for the small-tile shapes
gemm_ts_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TsParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  uint64_t* bars = (uint64_t*)(smem + (size_t)P.stages * P.stage_bytes);
  const uint32_t bar0 = smem_u32(bars);
  const int S = P.stages, NS = P.nslot;
  // barriers: full[S] (TMA bytes), a_full[S] (copy -> MMA), empty[S] (MMA retired -> producer), t_full[NS], t_empty[NS]
  const uint32_t full = bar0, a_full = full + 8 * MAX_S, empty = a_full + 8 * MAX_S, t_full = empty + 8 * MAX_S, t_empty = t_full + 8 * MAX_NS;
  uint32_t* tmem_word = (uint32_t*)(bars + 3 * MAX_S + 2 * MAX_NS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long G = gridDim.x, b = blockIdx.x;
  const long long n_local = (b < P.groups) ? (P.groups - b + G - 1) / G : 0;      // groups of P.grp tiles

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_b) : "memory");
    for (int i = 0; i < S; ++i) { mbar_init(full + 8 * i, 1); mbar_init(a_full + 8 * i, 4); mbar_init(empty + 8 * i, 1); }
    for (int i = 0; i < NS; ++i) { mbar_init(t_full + 8 * i, 1); mbar_init(t_empty + 8 * i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_word)), "r"((uint32_t)P.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_word;
  const uint32_t tmem_a = tmem_base + (uint32_t)P.a_col0;

  // producer and MMA warps loop with all lanes (warp-uniform values -> uniform registers); one elected lane issues (see gemm_tc.cu)
  if (warp == 0) {
    // ===================================== TMA producer =====================================
    {
      uint32_t leader;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
      int stage = 0; uint32_t phase = 0;
      for (long long i = 0; i < n_local; ++i) {
        const long long t = (b + i * G) * P.grp;                       // first tile of the group; the boxes span P.grp tiles
        for (unsigned long long r = 0; r < P.br; ++r) {
          for (int kc = 0; kc < P.kchunks; ++kc) {
            if (leader) mbar_wait(empty + 8 * stage, phase ^ 1);          // one lane polls, the others park at the warp barrier
            __syncwarp();
            const uint32_t sa = smem_base + stage * P.stage_bytes, sb = sa + P.a_bytes;
            if (leader) {
              if (P.skip & 8) mbar_arrive(full + 8 * stage);
              else {
                mbar_expect_tx(full + 8 * stage, (uint32_t)P.tx_bytes);         // rows / words beyond the matrix are zero-filled and counted
                tma_load_4d(sa, &map_a, full + 8 * stage, 0, kc * 32, (int)r, (int)t);               // m words x 32 word-rows
                tma_load_4d(sb, &map_b, full + 8 * stage, kc * P.kc_elems, 0, (int)r, (int)t);       // 128 bytes of k x np rows
              }
            }
            __syncwarp();
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    {
      uint32_t leader;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
      int stage = 0; uint32_t phase = 0;
      for (long long i = 0; i < n_local; ++i) {
        const int slot = (int)(i % NS);
        if (leader) mbar_wait(t_empty + 8 * slot, (uint32_t)(((i / NS) & 1) ^ 1));
        __syncwarp();
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(slot * P.slot_cols);
        uint32_t accumulate = 0;
        for (unsigned long long r = 0; r < P.br; ++r) {
          for (int kc = 0; kc < P.kchunks; ++kc) {
            if (leader) {
              mbar_wait(full + 8 * stage, phase);            // B bytes have landed (the copy warps waited on it too)
              mbar_wait(a_full + 8 * stage, phase);          // this stage's A words are in tensor memory
            }
            __syncwarp();
            tc_fence_after();
            const uint32_t sb = smem_base + stage * P.stage_bytes + P.a_bytes;
            const int krem = P.k - kc * P.kc_elems;
            const int ksteps = (krem >= P.kc_elems) ? 4 : (krem + P.kinst - 1) / P.kinst;
            const uint32_t a_slot = tmem_a + (uint32_t)stage * 32u;
            // B descriptor: constant high word; low word = (address >> 4) | 1 << 16; a k-step is 32 bytes inside the swizzled row
            const uint32_t b_lo = ((sb & 0x3FFFFu) >> 4) | (1u << 16), b_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
            if (leader && !(P.skip & 2)) {
              if (P.is_i8) { for (int ks = 0; ks < ksteps; ++ks) umma_i8_ts(d_tmem, a_slot + (uint32_t)ks * 8u, ((uint64_t)b_hi << 32) | (b_lo + (uint32_t)ks * 2u), P.idesc, ks == 0 ? accumulate : 1u); }
              else { for (int ks = 0; ks < ksteps; ++ks) umma_f16_ts(d_tmem, a_slot + (uint32_t)ks * 8u, ((uint64_t)b_hi << 32) | (b_lo + (uint32_t)ks * 2u), P.idesc, ks == 0 ? accumulate : 1u); }
            }
            accumulate = 1;
            if (leader) umma_commit(empty + 8 * stage);      // stage (shared memory AND its A columns) reusable once these retire
            __syncwarp();
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
        }
        if (leader) umma_commit(t_full + 8 * slot);
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ===================================== epilogue =========================================
    const int q = warp & 3;
    const int row = 32 * q + lane;
    for (long long i = 0; i < n_local; ++i) {
      const int slot = (int)(i % NS);
      mbar_wait(t_full + 8 * slot, (uint32_t)((i / NS) & 1));
      tc_fence_after();
      // row r of the instruction belongs to tile (group * grp + r / mp), row r % mp of it, and that tile's results sit in
      // accumulator columns (r / mp) * np ...: a warp covers 32 / mp tiles (one tcgen05.ld sequence each)
      const int tl = row / P.mp, ri = row - tl * P.mp;
      const long long tile = (b + i * G) * P.grp + tl;
      const bool valid = tl < P.grp && ri < P.m && tile < P.count && !(P.skip & 1);
      const long long ldcb = P.ldc * P.c_esz;
      char* crow = P.c + tile * P.tile_stride_c + (long long)ri * P.c_esz;
      const uint32_t taddr = tmem_base + (uint32_t)(slot * P.slot_cols) + ((uint32_t)(q * 32) << 16);
      const int wtl0 = (32 * q) / P.mp, nsub = (P.mp >= 32) ? 1 : 32 / P.mp;
      for (int sub = 0; sub < nsub; ++sub) {
        const int wtl = wtl0 + sub;                                     // warp-uniform tile index inside the group
        const bool last_sub = (sub == nsub - 1) || (wtl + 1 >= P.grp);
        if (wtl < P.grp) {
          for (int c0 = 0; c0 < P.np; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + (uint32_t)(wtl * P.np + c0), v);
            if (last_sub && c0 + 32 >= P.np) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(t_empty + 8 * slot); }
            if (valid && tl == wtl && c0 < P.n) xb_ep_store_chunk(P.ep_mode, P.beta0, v, crow + c0 * ldcb, ldcb, P.n - c0, P.scf);
          }
        } else if (sub == 0) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(t_empty + 8 * slot); }   // rows beyond the group's tiles
      }
    }
  } else {
    // ===================================== copy: raw VNNI words -> TMEM A =====================
    // this warp owns rows 32q .. 32q+31 (its TMEM lane quadrant); the raw stage is [32 word-rows][m words]: lanes read
    // consecutive words (conflict-free); rows >= m read a neighbouring stage's bytes -- they only feed accumulator rows
    // that are never stored
    const int q = warp & 3, row = 32 * q + lane;
    const int tl = row / P.mp, ri = row - tl * P.mp;
    const int rr = (tl < P.grp && ri < P.m) ? tl * 32 * P.m + ri : 0;      // the raw stage is [grp][32 word-rows][m words]
    int stage = 0; uint32_t phase = 0;
    for (long long i = 0; i < n_local; ++i) {
      for (unsigned long long r = 0; r < P.br; ++r) {
        for (int kc = 0; kc < P.kchunks; ++kc) {
          mbar_wait(full + 8 * stage, phase);
          const unsigned int* src = reinterpret_cast<const unsigned int*>(smem + (size_t)stage * P.stage_bytes) + rr;
          if (!(P.skip & 4)) {
          uint32_t w[32];
#pragma unroll
          for (int kv = 0; kv < 32; ++kv) w[kv] = src[(size_t)kv * P.m];
          tmem_st32(tmem_a + (uint32_t)stage * 32u + ((uint32_t)(q * 32) << 16), w);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_full + 8 * stage);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)P.tmem_cols) : "memory");
  }
}

int ts_env_int(const char* name, int fallback) { const char* e = getenv(name); return (e != nullptr && *e != 0) ? atoi(e) : fallback; }
unsigned long long g_ts_attr = 0ull;

}  // namespace

// ---- host side -------------------------------------------------------------------------------------------------------
extern "C" int xb_gemm_ts_supported(const xb_gemm_desc* d) {
  const unsigned int bad = LIBXSMM_GEMM_FLAG_TRANS_A | LIBXSMM_GEMM_FLAG_TRANS_B | LIBXSMM_GEMM_FLAG_VNNI_B | LIBXSMM_GEMM_FLAG_VNNI_C
                         | LIBXSMM_GEMM_FLAG_DECOMPRESS_A_VIA_BITMASK | LIBXSMM_GEMM_FLAG_INTLV_A_FORMAT;
  const int a8 = (d->ta == LIBXSMM_DATATYPE_I8 || d->ta == LIBXSMM_DATATYPE_U8), b8 = (d->tb == LIBXSMM_DATATYPE_I8 || d->tb == LIBXSMM_DATATYPE_U8);
  static int enabled = -1;
  if (enabled < 0) enabled = ts_env_int("LIBXSMM_B200_TS", 1) != 0 ? 1 : 0;
  if (!enabled || (d->flags & bad) != 0 || d->fuse_colbias != 0 || d->cp_op != 0) return 0;
  if (a8 && b8) {
    if (d->tcomp != LIBXSMM_DATATYPE_I32 || !(d->tc == LIBXSMM_DATATYPE_I32 || d->tc == LIBXSMM_DATATYPE_F32)) return 0;
    if ((d->flags & LIBXSMM_GEMM_FLAG_VNNI_A) == 0 && d->tc != LIBXSMM_DATATYPE_F32) return 0;   // flat int8 A: exact-order kernel (the F32-out path is always VNNI4)
    if ((d->k % 4) != 0 || (d->ldb % 16) != 0) return 0;
  } else if (d->ta == d->tb && (d->ta == LIBXSMM_DATATYPE_BF16 || d->ta == LIBXSMM_DATATYPE_F16)) {
    if ((d->flags & LIBXSMM_GEMM_FLAG_VNNI_A) == 0 || d->tcomp != LIBXSMM_DATATYPE_F32) return 0;
    if (d->ta == LIBXSMM_DATATYPE_BF16 && !(d->tc == LIBXSMM_DATATYPE_F32 || d->tc == LIBXSMM_DATATYPE_BF16)) return 0;
    if (d->ta == LIBXSMM_DATATYPE_F16 && !(d->tc == LIBXSMM_DATATYPE_F32 || d->tc == LIBXSMM_DATATYPE_F16)) return 0;
    if ((d->k % 2) != 0 || (d->ldb % 8) != 0) return 0;
  } else return 0;
  if (d->m < 4 || d->m > 128 || (d->m % 4) != 0 || (d->lda % 4) != 0 || d->n < 1 || d->n > 128 || d->k > 8192) return 0;
  if (!(d->br_type == 0 || d->br_type == 3)) return 0;
  if (d->br_type == 3 && ((d->br_stride_a % 16) != 0 || (d->br_stride_b % 16) != 0 || d->br_stride_a <= 0 || d->br_stride_b <= 0)) return 0;
  return 1;
}

extern "C" int xb_gemm_ts_launch(const xb_gemm_launch* L) {
  const xb_gemm_desc& d = L->d;
  const char* a = (const char*)L->a; const char* b = (const char*)L->b; char* c = (char*)L->c;
  long long sa = L->tile_stride_a, sb = L->tile_stride_b, sc = L->tile_stride_c;
  unsigned long long br = L->br;
  if (L->recs != nullptr) return xb_gemm_simt_launch(L);
  if (a == nullptr && c == nullptr) { a = (const char*)L->one.a; b = (const char*)L->one.b; c = (char*)L->one.c; br = L->one.br; sa = sb = sc = 0; }
  if (d.br_type == 0) br = 1;
  const bool aligned = (((uintptr_t)a | (uintptr_t)b) & 15) == 0 && (sa % 16) == 0 && (sb % 16) == 0 && (L->count == 1 || (sa > 0 && sb > 0));
  xb_encode_tiled_fn enc = xb_tma_encoder();
  if (br == 0 || !aligned || L->count <= 0 || br > 0x7fffffffull || L->count > 0x7fffffffll || enc == nullptr) return xb_gemm_simt_launch(L);

  const int is_i8 = (d.ta == LIBXSMM_DATATYPE_I8 || d.ta == LIBXSMM_DATATYPE_U8);
  const int v = is_i8 ? 4 : 2, es = is_i8 ? 1 : 2;
  TsParams P; memset(&P, 0, sizeof(P));
  P.m = d.m; P.n = d.n; P.k = d.k; P.np = (d.n + 15) & ~15;
  P.kc_elems = 128 / es; P.kinst = is_i8 ? 32 : 16; P.kchunks = (d.k + P.kc_elems - 1) / P.kc_elems;
  // small tiles: grp tiles at constant stride share one instruction as a block-diagonal product -- their A rows are stacked
  // (tile j at rows j*mp), their B columns put side by side (tile j at columns j*np); the off-diagonal blocks are computed and
  // ignored. One TMA box per operand fetches all grp tiles (4th box dimension), one barrier round trip serves grp tiles.
  P.grp = 1; P.mp = 128;
  if (d.m <= 32 && L->count > 1 && ts_env_int("LIBXSMM_B200_TS_PACK", 1) != 0) {
    P.mp = d.m <= 8 ? 8 : (d.m <= 16 ? 16 : 32);
    P.grp = 128 / P.mp; if (P.grp > 128 / P.np) P.grp = 128 / P.np;
    // 96 accumulator columns per slot leave room for two CTAs per SM (2 x 96 + 2 x 32 TMEM columns each): measured better than one
    // CTA with full 128-column groups (int8 32^3: 0.075 -> 0.066 ms)
    if (P.grp * P.np > 96 && 96 / P.np >= 2) P.grp = 96 / P.np;
    { const int ge = ts_env_int("LIBXSMM_B200_TS_GRP", 0); if (ge >= 1 && ge < P.grp) P.grp = ge; }
    if (P.grp < 2) { P.grp = 1; P.mp = 128; }
  }
  P.groups = (L->count + P.grp - 1) / P.grp;
  P.a_bytes = P.grp * d.m * 128;                               // grp x (m words x 32 word-rows)
  P.tx_bytes = P.a_bytes + P.grp * P.np * 128;
  P.stage_bytes = ((P.a_bytes + 1023) & ~1023) + P.grp * P.np * 128;   // B starts 1024-byte aligned (SWIZZLE_128B atom)
  P.a_bytes = (P.a_bytes + 1023) & ~1023;
  P.slot_cols = (P.grp * P.np + 31) & ~31;
  const long long loads_per_tile = (long long)P.kchunks * (long long)br;
  int ctas = ts_env_int("LIBXSMM_B200_TS_CTAS", loads_per_tile <= 2 ? 4 : 2);
  if (ctas < 1) ctas = 1; if (ctas > 4) ctas = 4;
  auto tmem_for = [](int c) { return c == 1 ? 512 : (c == 2 ? 256 : 128); };
  // TMEM per CTA: 2 accumulator slots + >= 2 A slots of 32 columns; shared memory: >= 2 stages
  while (ctas > 1 && (2 * P.slot_cols + 2 * 32 > tmem_for(ctas) || 2 * P.stage_bytes + 2048 > (224 * 1024) / ctas)) --ctas;
  P.tmem_cols = tmem_for(ctas);
  P.nslot = 2; if (P.nslot * P.slot_cols + 64 > P.tmem_cols) P.nslot = 1;
  int s_tmem = (P.tmem_cols - P.nslot * P.slot_cols) / 32, s_smem = ((224 * 1024) / ctas - 2048) / P.stage_bytes;
  P.stages = s_tmem < s_smem ? s_tmem : s_smem; if (P.stages > MAX_S) P.stages = MAX_S;
  { const int st = ts_env_int("LIBXSMM_B200_TS_STAGES", P.stages); if (st >= 2 && st <= P.stages) P.stages = st; }
  if (P.stages < 2) return xb_gemm_simt_launch(L);
  P.a_col0 = P.nslot * P.slot_cols;
  P.br = br; P.count = L->count; P.c = c; P.tile_stride_c = sc; P.ldc = d.ldc;
#if defined(XB_DIAG)   /* ablation switch of the profiling sessions: never in a release build (it makes the kernel skip work) */
  P.skip = ts_env_int("LIBXSMM_B200_TS_SKIP", 0);
#else
  P.skip = 0;
#endif
  P.ep_mode = xb_ep_mode(d.ta, d.tc, &P.c_esz);
  P.c_type = d.tc; P.a_type = d.ta; P.beta0 = (d.flags & LIBXSMM_GEMM_FLAG_BETA_0) ? 1 : 0; P.is_i8 = is_i8; P.scf = L->one.scf;
  if (is_i8) {
    const uint32_t fa = (d.ta == LIBXSMM_DATATYPE_I8) ? 1u : 0u, fb = (d.tb == LIBXSMM_DATATYPE_I8) ? 1u : 0u;
    P.idesc = (2u << 4) | (fa << 7) | (fb << 10) | ((uint32_t)((P.grp * P.np) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  } else {
    const uint32_t fmt = (d.ta == LIBXSMM_DATATYPE_BF16) ? 1u : 0u;
    P.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)((P.grp * P.np) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  }

  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const size_t ext_a = ((size_t)(d.k / v - 1) * d.lda + d.m) * 4, ext_b = ((size_t)(d.n - 1) * d.ldb + d.k) * es;
  const cuuint64_t pad_a = (ext_a + 15) & ~(size_t)15, pad_b = (ext_b + 15) & ~(size_t)15;
  CUtensorMap map_a, map_b;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)d.m, (cuuint64_t)(d.k / v), (cuuint64_t)br, (cuuint64_t)L->count};
    const cuuint64_t strides[3] = {(cuuint64_t)d.lda * 4, (d.br_type == 3) ? (cuuint64_t)d.br_stride_a : pad_a, (L->count > 1) ? (cuuint64_t)sa : pad_a};
    const cuuint32_t box[4] = {(cuuint32_t)d.m, 32, 1, (cuuint32_t)P.grp};
    if (CUDA_SUCCESS != enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, (void*)a, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return xb_gemm_simt_launch(L);
  }
  {
    const CUtensorMapDataType dt = is_i8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : (d.ta == LIBXSMM_DATATYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    const cuuint64_t dims[4] = {(cuuint64_t)d.k, (cuuint64_t)d.n, (cuuint64_t)br, (cuuint64_t)L->count};
    const cuuint64_t strides[3] = {(cuuint64_t)d.ldb * es, (d.br_type == 3) ? (cuuint64_t)d.br_stride_b : pad_b, (L->count > 1) ? (cuuint64_t)sb : pad_b};
    const cuuint32_t box[4] = {(cuuint32_t)P.kc_elems, (cuuint32_t)P.np, 1, (cuuint32_t)P.grp};
    if (CUDA_SUCCESS != enc(&map_b, dt, 4, (void*)b, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return xb_gemm_simt_launch(L);
  }
  const size_t smem = (size_t)P.stages * P.stage_bytes + 1024 + (3 * MAX_S + 2 * MAX_NS) * 8 + 64;
  static int sms = 0;
  if (sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  long long grid = P.groups; if (grid > (long long)sms * ctas) grid = (long long)sms * ctas; if (grid < 1) grid = 1;
  if (xb_rt_first_use_on_device(&g_ts_attr)) cudaFuncSetAttribute(gemm_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  gemm_ts_kernel<<<(unsigned int)grid, kTsgThreads, smem, (cudaStream_t)xb_rt_stream()>>>(map_a, map_b, P);
  xb_rt_count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { xb_rt_note_error((int)e, "gemm_ts"); return (int)e; }
  return 0;
}
