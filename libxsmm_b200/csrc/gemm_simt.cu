// libxsmm_b200 -- exact-order dense GEMM/BRGEMM on CUDA cores (sm_100a).
//
// This is the "every datatype, every layout flag" kernel: one CTA per tile, one thread per C element,
// and for each element the SAME sequence of multiply/add operations the reference's C kernel performs
// (src/generator_gemm_reference_impl.c:821-2800, libxsmm_ref_matmul). Because the order and the
// rounding points are identical (separate multiply and add, no contraction), results are bit-identical
// to the reference for integer AND floating-point types. The tensor-core kernel (gemm_tc.cu) is the
// fast path for the shapes it supports; this kernel is what every other descriptor launches.
//
// Layout formulas (elements), from the reference (file above, lines cited per branch):
//   A flat    A[k*lda + m]            A trans   A[m*lda + k]
//   A VNNI-v  A[(k/v)*lda*v + m*v + k%v]   (v = 2 for 16-bit, 4 for 8-bit; x86 pack factors)
//   B flat    B[n*ldb + k]            B trans   B[k*ldb + n]       B VNNI-T  B[(k/v)*ldb*v + n*v + k%v]
//   C         C[n*ldc + m]
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>
#include "xb_internal.h"
#include "xb_device.cuh"

namespace {

enum { P_F64 = 0, P_F32, P_I16, P_I8_I32, P_I8_F32, P_F16_F16, P_F16_F32, P_BF16_F32, P_BF16_BF16, P_I4_I32, P_BITMAP, P_FP8, P_NONE };

__host__ __device__ inline int xb_path_of(const xb_gemm_desc& d) {
  const int a = d.ta, b = d.tb, c = d.tc, comp = d.tcomp;
  const bool a8 = (a == LIBXSMM_DATATYPE_I8 || a == LIBXSMM_DATATYPE_U8);
  const bool b8 = (b == LIBXSMM_DATATYPE_I8 || b == LIBXSMM_DATATYPE_U8);
  if ((d.flags & LIBXSMM_GEMM_FLAG_DECOMPRESS_A_VIA_BITMASK) != 0) {   // bitmap-compressed A (reference :857-948): float operands, no batch reduce
    const bool fa = (a == LIBXSMM_DATATYPE_F32 || a == LIBXSMM_DATATYPE_BF16 || a == LIBXSMM_DATATYPE_F16);
    const bool fb = (b == LIBXSMM_DATATYPE_F32 || b == LIBXSMM_DATATYPE_BF16 || b == LIBXSMM_DATATYPE_F16);
    const bool fc = (c == LIBXSMM_DATATYPE_F32 || c == LIBXSMM_DATATYPE_BF16 || c == LIBXSMM_DATATYPE_F16);
    const int kb = (a == LIBXSMM_DATATYPE_F32) ? 1 : ((b == LIBXSMM_DATATYPE_F32) ? 1 : 2);
    return (fa && fb && fc && d.br_type == 0 && (d.k % kb) == 0 && d.fuse_colbias == 0 && d.cp_op == 0) ? P_BITMAP : P_NONE;
  }
  if (a == LIBXSMM_DATATYPE_I4X2 || a == LIBXSMM_DATATYPE_U4X2) {       // 4-bit A with zero points x 8-bit B -> I32 (reference :1273-1321)
    const unsigned int need = LIBXSMM_GEMM_FLAG_VNNI_A | LIBXSMM_GEMM_FLAG_INTLV_A_FORMAT;
    return ((d.flags & need) == need && b8 && c == LIBXSMM_DATATYPE_I32 && comp == LIBXSMM_DATATYPE_I32 && (d.k % 8) == 0
            && (d.br_type == 0 || d.br_type == 3)) ? P_I4_I32 : P_NONE;
  }
  if (a == LIBXSMM_DATATYPE_F64 && b == a && c == a && comp == a) return P_F64;
  if ((a == LIBXSMM_DATATYPE_F32 || a == LIBXSMM_DATATYPE_BF32) && (b == LIBXSMM_DATATYPE_F32 || b == LIBXSMM_DATATYPE_BF32)
      && c == LIBXSMM_DATATYPE_F32 && comp == LIBXSMM_DATATYPE_F32) return P_F32;
  if (a == LIBXSMM_DATATYPE_I16 && b == a && c == LIBXSMM_DATATYPE_I32 && comp == LIBXSMM_DATATYPE_I32) return P_I16;
  if (a8 && b8 && c == LIBXSMM_DATATYPE_I32 && comp == LIBXSMM_DATATYPE_I32) return P_I8_I32;
  if (a8 && b8 && c == LIBXSMM_DATATYPE_F32 && comp == LIBXSMM_DATATYPE_I32) return P_I8_F32;
  const bool f16comp = (comp == LIBXSMM_DATATYPE_F16 || comp == LIBXSMM_DATATYPE_F32 || comp == LIBXSMM_DATATYPE_IMPLICIT);
  if (a == LIBXSMM_DATATYPE_F16 && b == a && c == LIBXSMM_DATATYPE_F16 && f16comp) return P_F16_F16;
  if (a == LIBXSMM_DATATYPE_F16 && b == a && c == LIBXSMM_DATATYPE_F32 && f16comp) return P_F16_F32;
  if (a == LIBXSMM_DATATYPE_BF16 && b == a && c == LIBXSMM_DATATYPE_F32 && comp == LIBXSMM_DATATYPE_F32) return P_BF16_F32;
  if (a == LIBXSMM_DATATYPE_BF16 && b == a && c == LIBXSMM_DATATYPE_BF16 && comp == LIBXSMM_DATATYPE_F32) return P_BF16_BF16;
  // 8-bit float A (reference :2171-2366 with a bf16 B, :2420-2630 with B and C of A's type or f32): f32 accumulate, one rounding
  if ((a == LIBXSMM_DATATYPE_BF8 || a == LIBXSMM_DATATYPE_HF8) && comp == LIBXSMM_DATATYPE_F32 && d.fuse_colbias == 0 && d.cp_op == 0
      && (d.flags & (LIBXSMM_GEMM_FLAG_VNNI_C | LIBXSMM_GEMM_FLAG_VNNI_B)) == 0) {
    if (b == a && (c == LIBXSMM_DATATYPE_F32 || c == a)) return P_FP8;
    if (b == LIBXSMM_DATATYPE_BF16 && (c == LIBXSMM_DATATYPE_F32 || c == LIBXSMM_DATATYPE_BF16)) return P_FP8;
  }
  return P_NONE;
}

struct TileCtx {
  const char* a0; const char* b0; char* c;       // tile bases
  const void* const* a_addr; const void* const* b_addr;   // address mode
  const long long* a_offs; const long long* b_offs;       // offset mode
  unsigned long long br;
  float scf;
  const void* colbias; unsigned char* relu_mask;          // fused form (libxsmm_dispatch_brgemm_ext)
  const unsigned char* a_q;                                // int4: zero points; bitmap-compressed A: the bitmap
};

__device__ inline void resolve_tile(const xb_gemm_launch& L, long long t, TileCtx& x) {
  xb_gemm_rec r;
  if (L.recs != nullptr) r = L.recs[t];
  else if (L.a == nullptr && L.c == nullptr) r = L.one;
  else {
    r.a = (const char*)L.a + t * L.tile_stride_a; r.b = (const char*)L.b + t * L.tile_stride_b;
    r.c = (char*)L.c + t * L.tile_stride_c; r.a_aux = nullptr; r.b_aux = nullptr; r.br = L.br; r.scf = L.one.scf;
    r.d = L.one.d; r.c_aux = nullptr;                      // a shared bias column; masks only exist per call
    r.a_q = L.one.a_q;
    if (L.d.br_type == 2) { r.a_aux = L.one.a_aux; r.b_aux = L.one.b_aux; }
  }
  x.a0 = (const char*)r.a; x.b0 = (const char*)r.b; x.c = (char*)r.c;
  x.a_addr = (const void* const*)r.a; x.b_addr = (const void* const*)r.b;
  x.a_offs = (const long long*)r.a_aux; x.b_offs = (const long long*)r.b_aux;
  x.br = (L.d.br_type == 0) ? 1ull : r.br; x.scf = r.scf;
  x.colbias = r.d; x.relu_mask = (unsigned char*)r.c_aux; x.a_q = (const unsigned char*)r.a_q;
}

// base pointers of the r-th batch-reduce operand pair; mirrors libxsmm_calculate_brgemm_offsets
// (generator_gemm_reference_impl.c:178-197): byte offsets/strides are truncated to whole elements.
__device__ inline void br_ptrs(const xb_gemm_desc& d, const TileCtx& x, unsigned long long r, int tsa, int tsb,
                               const char*& pa, const char*& pb) {
  switch (d.br_type) {
    case 1: pa = (const char*)x.a_addr[r]; pb = (const char*)x.b_addr[r]; break;
    case 2: pa = x.a0 + (x.a_offs[r] / tsa) * tsa; pb = x.b0 + (x.b_offs[r] / tsb) * tsb; break;
    case 3: pa = x.a0 + (long long)r * ((d.br_stride_a / tsa) * tsa); pb = x.b0 + (long long)r * ((d.br_stride_b / tsb) * tsb); break;
    default: pa = x.a0; pb = x.b0;
  }
}

template <typename T> __device__ inline T ldg_as(const char* base, long long idx) {
  return reinterpret_cast<const T*>(base)[idx];
}

// ---- fused form: column-bias pre-op, ReLU (+bitmask) / sigmoid post-op (reference :255-372) -----------------------------
// The reference builds an f32 image of C (bias column broadcast, + old C when beta = 1), lets the GEMM accumulate into it
// with beta = 1 and C type F32, applies the post-op and rounds ONCE into C. Per element that is: seed -> same loop as the
// F32-output variant of the precision path -> activation -> one conversion. The pre-ops are mateltwise kernels, hence their
// bf16 load (flushes bf16 subnormals).
__device__ __forceinline__ float fuse_ld(const void* p, long long i, int t) {
  if (t == LIBXSMM_DATATYPE_F32) return ((const float*)p)[i];
  if (t == LIBXSMM_DATATYPE_BF16) { unsigned short h = ((const unsigned short*)p)[i]; if ((h & 0x7f80) == 0) h &= 0x8000; return xb_bf16_to_f32(h); }
  return xb_f16_to_f32(((const unsigned short*)p)[i]);
}
__device__ __forceinline__ void fuse_st(void* p, long long i, int t, float v) {
  if (t == LIBXSMM_DATATYPE_F32) ((float*)p)[i] = v;
  else if (t == LIBXSMM_DATATYPE_BF16) ((unsigned short*)p)[i] = xb_f32_to_bf16_rne(v);
  else ((unsigned short*)p)[i] = xb_f32_to_f16(v);
}

__global__ void __launch_bounds__(256) gemm_simt_fused_kernel(const xb_gemm_launch L, const int path) {
  const xb_gemm_desc& d = L.d;
  const int m = d.m, n = d.n, k = d.k;
  const long long lda = d.lda, ldb = d.ldb, ldc = d.ldc;
  const bool trans_a = (d.flags & LIBXSMM_GEMM_FLAG_TRANS_A) != 0, trans_b = (d.flags & LIBXSMM_GEMM_FLAG_TRANS_B) != 0;
  const bool vnni_a = (d.flags & LIBXSMM_GEMM_FLAG_VNNI_A) != 0, vnni_b = (d.flags & LIBXSMM_GEMM_FLAG_VNNI_B) != 0;
  const bool ua = (d.ta == LIBXSMM_DATATYPE_U8), ub = (d.tb == LIBXSMM_DATATYPE_U8);
  const int tsa = xb_dev_typesize(d.ta), tsb = xb_dev_typesize(d.tb);
  const bool bias = d.fuse_colbias != 0, relu = d.cp_op == LIBXSMM_MELTW_TYPE_UNARY_RELU, sigm = d.cp_op == LIBXSMM_MELTW_TYPE_UNARY_SIGMOID;
  const bool bitm = relu && (d.cp_flags & LIBXSMM_MELTW_FLAG_UNARY_BITMASK_2BYTEMULT) != 0;
  const bool beta0 = (d.flags & LIBXSMM_GEMM_FLAG_BETA_0) != 0, beta0_eff = beta0 && !bias;
  const int lane = threadIdx.x & 31, chunks = (m + 31) / 32;
  const long long mask_ld = ((ldc + 15) / 16) * 16;
  for (long long t = blockIdx.x; t < L.count; t += gridDim.x) {
    TileCtx x; resolve_tile(L, t, x);
    // a warp owns 32 consecutive rows of one column, so that the ReLU bitmask is one ballot per warp
    for (int w = threadIdx.x >> 5; w < chunks * n; w += blockDim.x >> 5) {
      const int j = w / chunks, i0 = (w % chunks) * 32, i = i0 + lane;
      const bool act = i < m;
      const long long ci = (long long)j * ldc + i;
      float y = 0.0f, seed = 0.0f;
      if (act) {
        if (bias) { const float bv = fuse_ld(x.colbias, i, d.tc); seed = beta0 ? bv : __fadd_rn(bv, fuse_ld(x.c, ci, d.tc)); }
        else if (!beta0) seed = (d.tc == LIBXSMM_DATATYPE_F32) ? ((const float*)x.c)[ci] : fuse_ld(x.c, ci, d.tc);
        float acc = 0.0f;
        switch (path) {
          case P_F32: {
            const bool cvt = (d.ta == LIBXSMM_DATATYPE_BF32);
            acc = beta0_eff ? 0.0f : seed;
            for (unsigned long long r = 0; r < x.br; ++r) {
              const char *pa, *pb; br_ptrs(d, x, r, 4, 4, pa, pb);
              for (int s2 = 0; s2 < k; ++s2) {
                float av = ldg_as<float>(pa, trans_a ? (i * lda + s2) : (s2 * lda + i));
                float bv = ldg_as<float>(pb, trans_b ? (s2 * ldb + j) : (j * ldb + s2));
                if (cvt) { av = xb_bf16_to_f32(xb_f32_to_bf16_rne(av)); bv = xb_bf16_to_f32(xb_f32_to_bf16_rne(bv)); }
                acc = __fadd_rn(acc, __fmul_rn(av, bv));
              }
            }
          } break;
          case P_I8_F32: {
            unsigned int ia = 0u;
            for (unsigned long long r = 0; r < x.br; ++r) {
              const char *pa, *pb; br_ptrs(d, x, r, 1, 1, pa, pb);
              for (int s2 = 0; s2 < k / 4; ++s2) for (int k2 = 0; k2 < 4; ++k2) {
                const unsigned char ar = ldg_as<unsigned char>(pa, s2 * (lda * 4) + (long long)i * 4 + k2);
                const unsigned char brw = ldg_as<unsigned char>(pb, j * ldb + (long long)s2 * 4 + k2);
                ia += (unsigned int)((ua ? (int)ar : (int)(signed char)ar) * (ub ? (int)brw : (int)(signed char)brw));
              }
            }
            acc = __fmul_rn((float)(int)ia, x.scf);
            if (!beta0_eff) acc = __fadd_rn(acc, seed);
          } break;
          case P_F16_F16: case P_F16_F32: {
            const int kb = vnni_a ? 2 : 1;
            const bool round_each = (d.tcomp == LIBXSMM_DATATYPE_F16 || d.tcomp == LIBXSMM_DATATYPE_IMPLICIT);
            for (unsigned long long r = 0; r < x.br; ++r) {
              const char *pa, *pb; br_ptrs(d, x, r, 2, 2, pa, pb);
              for (int s2 = 0; s2 < k / kb; ++s2) for (int k2 = 0; k2 < kb; ++k2) {
                const float av = xb_f16_to_f32(ldg_as<unsigned short>(pa, s2 * (lda * kb) + (long long)i * kb + k2));
                const long long kk = (long long)s2 * kb + k2;
                const float bv = xb_f16_to_f32(ldg_as<unsigned short>(pb, trans_b ? (kk * ldb + j) : (j * ldb + kk)));
                acc = __fadd_rn(acc, __fmul_rn(av, bv));
                if (round_each) acc = xb_f16_to_f32(xb_f32_to_f16(acc));
              }
            }
            if (!beta0_eff) acc = __fadd_rn(acc, xb_f16_to_f32(xb_f32_to_f16(seed)));   // the F32-C variant rounds the old C through f16 (:2118-2124)
          } break;
          default: {   // P_BF16_F32 / P_BF16_BF16
            const int kb = vnni_a ? 2 : 1;
            acc = beta0_eff ? 0.0f : seed;
            for (unsigned long long r = 0; r < x.br; ++r) {
              const char *pa, *pb; br_ptrs(d, x, r, 2, 2, pa, pb);
              for (int s2 = 0; s2 < k / kb; ++s2) for (int k2 = kb - 1; k2 >= 0; --k2) {
                const long long kk = (long long)s2 * kb + k2;
                unsigned short ar = 0, brw = 0;
                if (!trans_a) ar = ldg_as<unsigned short>(pa, s2 * (lda * kb) + (long long)i * kb + k2);
                else if (!vnni_a) ar = ldg_as<unsigned short>(pa, i * lda + kk);
                if (trans_b && vnni_b) brw = ldg_as<unsigned short>(pb, (long long)j * kb + s2 * (ldb * kb) + k2);
                else if (trans_b) brw = ldg_as<unsigned short>(pb, kk * ldb + j);
                else if (!vnni_b) brw = ldg_as<unsigned short>(pb, j * ldb + kk);
                acc = __fadd_rn(acc, __fmul_rn(xb_bf16_to_f32(ar), xb_bf16_to_f32(brw)));
              }
            }
          } break;
        }
        y = relu ? ((acc <= 0.0f) ? 0.0f : acc) : (sigm ? (tanhf(acc / 2.0f) + 1.0f) / 2.0f : acc);
        fuse_st(x.c, ci, d.tc, y);
        seed = acc;
      }
      if (bitm) {
        const unsigned int word = __ballot_sync(0xffffffffu, act && !(seed <= 0.0f));
        if (lane < 4) {
          const int ib = i0 + lane * 8;
          if (ib < m) {
            unsigned char* dst = x.relu_mask + ib / 8 + (long long)j * (mask_ld / 8);
            const unsigned int valid = (m - ib >= 8) ? 0xffu : ((1u << (m - ib)) - 1u);
            const unsigned int nb = (word >> (lane * 8)) & 0xffu;
            *dst = (unsigned char)((valid == 0xffu) ? nb : ((*dst & ~valid) | (nb & valid)));
          }
        }
      }
    }
  }
  (void)tsa; (void)tsb;
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(const xb_gemm_launch L, const int path) {
  const xb_gemm_desc& d = L.d;
  const int m = d.m, n = d.n, k = d.k;
  const long long lda = d.lda, ldb = d.ldb, ldc = d.ldc;
  const bool beta0 = (d.flags & LIBXSMM_GEMM_FLAG_BETA_0) != 0;
  const bool trans_a = (d.flags & LIBXSMM_GEMM_FLAG_TRANS_A) != 0;
  const bool trans_b = (d.flags & LIBXSMM_GEMM_FLAG_TRANS_B) != 0;
  const bool vnni_a = (d.flags & LIBXSMM_GEMM_FLAG_VNNI_A) != 0;
  const bool vnni_b = (d.flags & LIBXSMM_GEMM_FLAG_VNNI_B) != 0;
  const bool ua = (d.ta == LIBXSMM_DATATYPE_U8), ub = (d.tb == LIBXSMM_DATATYPE_U8);
  const int tsa = xb_dev_typesize(d.ta), tsb = xb_dev_typesize(d.tb);

  for (long long t = blockIdx.x; t < L.count; t += gridDim.x) {
    TileCtx x; resolve_tile(L, t, x);
    for (int e = threadIdx.x; e < m * n; e += blockDim.x) {
      const int i = e % m, j = e / m;
      const long long ci = (long long)j * ldc + i;
      switch (path) {
        case P_F64: {   // reference :1322-1358, accumulates in place in C
          double acc = beta0 ? 0.0 : ldg_as<double>(x.c, ci);
          for (unsigned long long r = 0; r < x.br; ++r) {
            const char *pa, *pb; br_ptrs(d, x, r, 8, 8, pa, pb);
            for (int s = 0; s < k; ++s) {
              const double av = ldg_as<double>(pa, trans_a ? (i * lda + s) : (s * lda + i));
              const double bv = ldg_as<double>(pb, trans_b ? (s * ldb + j) : (j * ldb + s));
              acc = __dadd_rn(acc, __dmul_rn(av, bv));
            }
          }
          reinterpret_cast<double*>(x.c)[ci] = acc;
        } break;
        case P_F32: {   // reference :1359-1426 (BF32: operands rounded to bf16 first)
          const bool cvt = (d.ta == LIBXSMM_DATATYPE_BF32);
          float acc = beta0 ? 0.0f : ldg_as<float>(x.c, ci);
          for (unsigned long long r = 0; r < x.br; ++r) {
            const char *pa, *pb; br_ptrs(d, x, r, 4, 4, pa, pb);
            for (int s = 0; s < k; ++s) {
              float av = ldg_as<float>(pa, trans_a ? (i * lda + s) : (s * lda + i));
              float bv = ldg_as<float>(pb, trans_b ? (s * ldb + j) : (j * ldb + s));
              if (cvt) { av = xb_bf16_to_f32(xb_f32_to_bf16_rne(av)); bv = xb_bf16_to_f32(xb_f32_to_bf16_rne(bv)); }
              acc = __fadd_rn(acc, __fmul_rn(av, bv));
            }
          }
          reinterpret_cast<float*>(x.c)[ci] = acc;
        } break;
        case P_I16: {   // reference :1427-1451 (trans flags ignored, VNNI2 A optional)
          const int kb = vnni_a ? 2 : 1;
          int acc = beta0 ? 0 : ldg_as<int>(x.c, ci);
          for (unsigned long long r = 0; r < x.br; ++r) {
            const char *pa, *pb; br_ptrs(d, x, r, 2, 2, pa, pb);
            for (int s = 0; s < k / kb; ++s) for (int k2 = 0; k2 < kb; ++k2) {
              const int av = ldg_as<short>(pa, s * (lda * kb) + (long long)i * kb + k2);
              const int bv = ldg_as<short>(pb, j * ldb + (long long)s * kb + k2);
              acc += av * bv;
            }
          }
          reinterpret_cast<int*>(x.c)[ci] = acc;
        } break;
        case P_I8_I32: case P_I8_F32: {   // reference :1452-1683 (four sign combinations)
          const int kb = (path == P_I8_F32) ? 4 : (vnni_a ? 4 : 1);
          unsigned int acc = (path == P_I8_I32 && !beta0) ? (unsigned int)ldg_as<int>(x.c, ci) : 0u;
          for (unsigned long long r = 0; r < x.br; ++r) {
            const char *pa, *pb; br_ptrs(d, x, r, 1, 1, pa, pb);
            for (int s = 0; s < k / kb; ++s) for (int k2 = 0; k2 < kb; ++k2) {
              const unsigned char ar = ldg_as<unsigned char>(pa, s * (lda * kb) + (long long)i * kb + k2);
              const unsigned char brw = ldg_as<unsigned char>(pb, j * ldb + (long long)s * kb + k2);
              const int av = ua ? (int)ar : (int)(signed char)ar;
              const int bv = ub ? (int)brw : (int)(signed char)brw;
              acc += (unsigned int)(av * bv);      // wrap-around like the reference's int accumulator
            }
          }
          if (path == P_I8_I32) reinterpret_cast<int*>(x.c)[ci] = (int)acc;
          else {
            float f = __fmul_rn((float)(int)acc, x.scf);
            if (!beta0) f = __fadd_rn(f, ldg_as<float>(x.c, ci));
            reinterpret_cast<float*>(x.c)[ci] = f;
          }
        } break;
        case P_F16_F16: case P_F16_F32: {   // reference :2025-2126
          const int kb = vnni_a ? 2 : 1;
          // comp F16 (or IMPLICIT, resolved like an SPR host) rounds the accumulator to f16 per FMA
          const bool round_each = (d.tcomp == LIBXSMM_DATATYPE_F16 || d.tcomp == LIBXSMM_DATATYPE_IMPLICIT);
          float acc = 0.0f;
          for (unsigned long long r = 0; r < x.br; ++r) {
            const char *pa, *pb; br_ptrs(d, x, r, 2, 2, pa, pb);
            for (int s = 0; s < k / kb; ++s) for (int k2 = 0; k2 < kb; ++k2) {
              const float av = xb_f16_to_f32(ldg_as<unsigned short>(pa, s * (lda * kb) + (long long)i * kb + k2));
              const long long kk = (long long)s * kb + k2;
              const float bv = xb_f16_to_f32(ldg_as<unsigned short>(pb, trans_b ? (kk * ldb + j) : (j * ldb + kk)));
              acc = __fadd_rn(acc, __fmul_rn(av, bv));
              if (round_each) acc = xb_f16_to_f32(xb_f32_to_f16(acc));
            }
          }
          if (path == P_F16_F16) {
            if (!beta0) acc = __fadd_rn(acc, xb_f16_to_f32(ldg_as<unsigned short>(x.c, ci)));
            reinterpret_cast<unsigned short*>(x.c)[ci] = xb_f32_to_f16(acc);
          } else {
            if (!beta0) acc = __fadd_rn(acc, xb_f16_to_f32(xb_f32_to_f16(ldg_as<float>(x.c, ci))));
            reinterpret_cast<float*>(x.c)[ci] = acc;
          }
        } break;
        case P_BF16_F32: case P_BF16_BF16: {   // reference :2127-2170 and :2367-2419
          const int kb = vnni_a ? 2 : 1;
          float acc;
          if (path == P_BF16_F32) acc = beta0 ? 0.0f : ldg_as<float>(x.c, ci);
          else acc = beta0 ? 0.0f : xb_bf16_to_f32(ldg_as<unsigned short>(x.c, ci));
          for (unsigned long long r = 0; r < x.br; ++r) {
            const char *pa, *pb; br_ptrs(d, x, r, 2, 2, pa, pb);
            for (int s = 0; s < k / kb; ++s) for (int k2 = kb - 1; k2 >= 0; --k2) {   // high k of a pair first
              const long long kk = (long long)s * kb + k2;
              unsigned short ar = 0, brw = 0;
              if (!trans_a) ar = ldg_as<unsigned short>(pa, s * (lda * kb) + (long long)i * kb + k2);
              else if (!vnni_a) ar = ldg_as<unsigned short>(pa, i * lda + kk);
              if (trans_b && vnni_b) brw = ldg_as<unsigned short>(pb, (long long)j * kb + s * (ldb * kb) + k2);
              else if (trans_b) brw = ldg_as<unsigned short>(pb, kk * ldb + j);
              else if (!vnni_b) brw = ldg_as<unsigned short>(pb, j * ldb + kk);
              acc = __fadd_rn(acc, __fmul_rn(xb_bf16_to_f32(ar), xb_bf16_to_f32(brw)));
            }
          }
          if (path == P_BF16_F32) reinterpret_cast<float*>(x.c)[ci] = acc;
          else reinterpret_cast<unsigned short*>(x.c)[ci] = xb_f32_to_bf16_rne(acc);
        } break;
        case P_FP8: {   // reference :2420-2630 (B of A's 8-bit type: k ascending, VNNI factor 4) and :2171-2366 (bf16 B: pairs, high k first)
          const bool b16 = (d.tb == LIBXSMM_DATATYPE_BF16), hf = (d.ta == LIBXSMM_DATATYPE_HF8);
          const int kb = vnni_a ? (b16 ? 2 : 4) : 1;
          float acc = 0.0f;
          if (!beta0) {
            if (d.tc == LIBXSMM_DATATYPE_F32) acc = ldg_as<float>(x.c, ci);
            else if (d.tc == LIBXSMM_DATATYPE_BF16) acc = xb_bf16_to_f32(ldg_as<unsigned short>(x.c, ci));
            else acc = hf ? xb_hf8_to_f32(ldg_as<unsigned char>(x.c, ci)) : xb_bf8_to_f32(ldg_as<unsigned char>(x.c, ci));
          }
          for (unsigned long long r = 0; r < x.br; ++r) {
            const char *pa, *pb; br_ptrs(d, x, r, 1, b16 ? 2 : 1, pa, pb);
            for (int s = 0; s < k / kb; ++s) for (int q = 0; q < kb; ++q) {
              const int k2 = b16 ? (kb - 1 - q) : q;
              const long long kk = (long long)s * kb + k2;
              unsigned char ar = 0;
              if (!trans_a) ar = ldg_as<unsigned char>(pa, s * (lda * kb) + (long long)i * kb + k2);
              else if (!vnni_a) ar = ldg_as<unsigned char>(pa, i * lda + kk);
              float bv;
              if (b16) bv = xb_bf16_to_f32(trans_b ? ldg_as<unsigned short>(pb, kk * ldb + j) : ldg_as<unsigned short>(pb, j * ldb + kk));
              else { const unsigned char bw = trans_b ? ldg_as<unsigned char>(pb, kk * ldb + j) : ldg_as<unsigned char>(pb, j * ldb + kk); bv = hf ? xb_hf8_to_f32(bw) : xb_bf8_to_f32(bw); }
              acc = __fadd_rn(acc, __fmul_rn(hf ? xb_hf8_to_f32(ar) : xb_bf8_to_f32(ar), bv));
            }
          }
          if (d.tc == LIBXSMM_DATATYPE_F32) reinterpret_cast<float*>(x.c)[ci] = acc;
          else if (d.tc == LIBXSMM_DATATYPE_BF16) reinterpret_cast<unsigned short*>(x.c)[ci] = xb_f32_to_bf16_rne(acc);
          else reinterpret_cast<unsigned char*>(x.c)[ci] = hf ? xb_f32_to_hf8(acc) : xb_f32_to_bf8(acc);
        } break;
        case P_I4_I32: {   // reference :1273-1321; zero point subtracted in 8-bit arithmetic, B read as unsigned bytes
          unsigned int acc = beta0 ? 0u : (unsigned int)ldg_as<int>(x.c, ci);
          for (unsigned long long r = 0; r < x.br; ++r) {
            const char *pa, *pb; br_ptrs(d, x, r, 1, 1, pa, pb);
            const unsigned char z = (d.br_type == 3) ? x.a_q[((d.br_stride_a * 2) / k) * (long long)r + i] : x.a_q[i];
            for (int s = 0; s < k / 8; ++s) for (int q = 0; q < 4; ++q) {
              const unsigned char pk = ldg_as<unsigned char>(pa, s * lda * 4 + 4 * (long long)i + q);
              const int ev = (int)(signed char)((pk & 0x0f) - z), od = (int)(signed char)(((pk >> 4) & 0x0f) - z);
              acc += (unsigned int)(ev * (int)ldg_as<unsigned char>(pb, j * ldb + (long long)s * 8 + q));
              acc += (unsigned int)(od * (int)ldg_as<unsigned char>(pb, j * ldb + (long long)s * 8 + 4 + q));
            }
          }
          reinterpret_cast<int*>(x.c)[ci] = (int)acc;
        } break;
        default: break;
      }
    }
  }
}

// ---- bitmap-compressed A (DECOMPRESS_A_VIA_BITMASK, reference :857-948) ---------------------------------------------------
// A stores only the elements whose bit is set, in bit order; bit (s, i, k2) sits at position s*(m*kb) + i*kb + k2. The index
// of an element in the compressed array is the number of set bits in front of it: phase 1 scans the bitmap once (per-word
// population counts -> exclusive prefix in `prefix`), phase 2 is the exact-order element loop of the reference with
// idx = prefix[word] + popc(bits below). One CTA per call (the flag excludes batch reduce; single-tile launches only).
__device__ __forceinline__ unsigned int bitmap_word(const unsigned char* bm, long long w, long long nbytes) {
  unsigned int v = 0;
  for (int q = 0; q < 4; ++q) { const long long bi = w * 4 + q; if (bi < nbytes) v |= (unsigned int)bm[bi] << (8 * q); }
  return v;
}
__global__ void __launch_bounds__(1024) gemm_bitmap_kernel(const xb_gemm_launch L, unsigned int* __restrict__ prefix) {
  const xb_gemm_desc& d = L.d;
  const int m = d.m, n = d.n, k = d.k;
  const long long ldb = d.ldb, ldc = d.ldc;
  const bool beta0 = (d.flags & LIBXSMM_GEMM_FLAG_BETA_0) != 0;
  const int kb = (d.ta == LIBXSMM_DATATYPE_F32) ? 1 : ((d.tb == LIBXSMM_DATATYPE_F32) ? 1 : 2);
  const long long nbits = (long long)m * k, nbytes = (nbits + 7) / 8, nwords = (nbits + 31) / 32;
  const unsigned char* bm = (const unsigned char*)L.one.a_q;
  const char* a = (const char*)L.one.a; const char* b = (const char*)L.one.b; char* c = (char*)L.one.c;
  __shared__ unsigned int seg_total[1024];
  // phase 1: thread t owns a contiguous range of words
  const long long per = (nwords + blockDim.x - 1) / blockDim.x, w0 = (long long)threadIdx.x * per, w1 = (w0 + per < nwords) ? w0 + per : nwords;
  unsigned int run = 0;
  for (long long w = w0; w < w1; ++w) { prefix[w] = run; run += __popc(bitmap_word(bm, w, nbytes)); }
  seg_total[threadIdx.x] = run;
  __syncthreads();
  if (threadIdx.x == 0) { unsigned int acc = 0; for (unsigned int t = 0; t < blockDim.x; ++t) { const unsigned int v = seg_total[t]; seg_total[t] = acc; acc += v; } }
  __syncthreads();
  { const unsigned int base = seg_total[threadIdx.x]; for (long long w = w0; w < w1; ++w) prefix[w] += base; }
  __syncthreads();
  // phase 2
  for (int e = threadIdx.x; e < m * n; e += blockDim.x) {
    const int i = e % m, j = e / m;
    const long long ci = (long long)j * ldc + i;
    float acc;
    if (d.tc == LIBXSMM_DATATYPE_F32) acc = beta0 ? 0.0f : ((const float*)c)[ci];
    else acc = beta0 ? 0.0f : ((d.tc == LIBXSMM_DATATYPE_BF16) ? xb_bf16_to_f32(((const unsigned short*)c)[ci]) : xb_f16_to_f32(((const unsigned short*)c)[ci]));
    for (int s = 0; s < k / kb; ++s) for (int k2 = 0; k2 < kb; ++k2) {
      const long long bit = (long long)s * m * kb + (long long)i * kb + k2, w = bit >> 5;
      const unsigned int word = bitmap_word(bm, w, nbytes), sh = (unsigned int)(bit & 31);
      if ((word >> sh) & 1u) {
        const long long idx = (long long)prefix[w] + __popc(word & ((1u << sh) - 1u));
        const float av = (d.ta == LIBXSMM_DATATYPE_F32) ? ((const float*)a)[idx]
                       : ((d.ta == LIBXSMM_DATATYPE_BF16) ? xb_bf16_to_f32(((const unsigned short*)a)[idx]) : xb_f16_to_f32(((const unsigned short*)a)[idx]));
        const long long bi = j * ldb + (long long)s * kb + k2;
        const float bv = (d.tb == LIBXSMM_DATATYPE_F32) ? ((const float*)b)[bi]
                       : ((d.tb == LIBXSMM_DATATYPE_BF16) ? xb_bf16_to_f32(((const unsigned short*)b)[bi]) : xb_f16_to_f32(((const unsigned short*)b)[bi]));
        acc = __fadd_rn(acc, __fmul_rn(av, bv));
      }
    }
    if (d.tc == LIBXSMM_DATATYPE_F32) ((float*)c)[ci] = acc;
    else if (d.tc == LIBXSMM_DATATYPE_BF16) ((unsigned short*)c)[ci] = xb_f32_to_bf16_rne(acc);
    else ((unsigned short*)c)[ci] = xb_f32_to_f16(acc);
  }
}


// ---- 8-bit integer tiles: dp4a kernel ------------------------------------------------------------------------------
// Integer sums wrap modulo 2^32 and are therefore exact in ANY order: the int8 paths need not follow the reference's
// loop order to stay bit-identical (reference :1452-1683). One WARP per tile: the VNNI4 A words [k/4][m] and the
// k-contiguous B words [n][k/4] of a 64-wide k chunk are staged in shared memory with coalesced 4-byte loads, every
// lane keeps a TM x TN block of accumulators and issues one dp4a per (m, n, 4 k). HBM-bound by construction
// (m*k + k*n + 4*m*n bytes per tile); 8 warps per CTA and several CTAs per SM hide the load latency.
template <bool UA, bool UB> __device__ __forceinline__ unsigned int dp4a_x(unsigned int a, unsigned int b, unsigned int c) {
  unsigned int d;
  if (UA && UB) asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (UA && !UB) asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (!UA && UB) asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

constexpr int I8_WARPS = 16, I8_KC = 64;   // warps per CTA, k bytes per staged chunk

__device__ __forceinline__ void group_sync(int wpt, int group) {
  if (wpt == 1) __syncwarp();
  else asm volatile("bar.sync %0, %1;" :: "r"(group + 1), "r"(wpt * 32) : "memory");
}

// WPT warps share one tile: each owns one (LM*TM) x (LN*TN) block of C and keeps it in registers over the whole k loop;
// the group stages the operand chunk once. 16/WPT tiles are in flight per CTA.
template <int TM, int TN, bool UA, bool UB>
__global__ void __launch_bounds__(I8_WARPS * 32, 2) gemm_i8_kernel(const xb_gemm_launch L, const int to_f32, const int wpt, const int smem_words_per_group) {
  constexpr int LM = 8, LN = 4;                       // lanes along m and n
  extern __shared__ unsigned int smem_i8[];
  const xb_gemm_desc& d = L.d;
  const int m = d.m, n = d.n, kq_all = d.k >> 2;
  const long long lda = d.lda, ldb = d.ldb, ldc = d.ldc;
  const bool beta0 = (d.flags & LIBXSMM_GEMM_FLAG_BETA_0) != 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int groups = I8_WARPS / wpt, group = warp / wpt, wg = warp % wpt, gtid = wg * 32 + lane, gsize = wpt * 32;
  const int lm = lane % LM, ln = lane / LM;
  const int passes_m = (m + LM * TM - 1) / (LM * TM);
  const int m0 = (wg % passes_m) * (LM * TM), n0 = (wg / passes_m) * (LN * TN);   // this warp's block (may be empty: n0 >= n)
  const int kcq = (kq_all < I8_KC / 4) ? kq_all : I8_KC / 4;      // words of k per chunk
  // both panels are stored k-group-major ([q][m] and [q][n]) so that a lane's TM / TN operands are contiguous (128-bit shared
  // loads); row strides are 4 * odd words: 16-byte aligned, and the transposing B fill spreads over the banks
  const int ms = (((m + 3) >> 2) | 1) << 2, ns = (((n + 3) >> 2) | 1) << 2;
  unsigned int* sa = smem_i8 + (size_t)group * smem_words_per_group;   // [kcq][ms]
  unsigned int* sb = sa + (size_t)kcq * ms;                            // [kcq][ns]
  const long long ntiles_rounded = ((L.count + (long long)gridDim.x * groups - 1) / ((long long)gridDim.x * groups)) * ((long long)gridDim.x * groups);
  for (long long t = (long long)blockIdx.x * groups + group; t < ntiles_rounded; t += (long long)gridDim.x * groups) {
    const bool live = t < L.count;                    // dead iterations only keep the group barriers balanced
    TileCtx x; if (live) resolve_tile(L, t, x); else { x.br = 0; }
    unsigned int acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0u;
    if (live) for (unsigned long long r = 0; r < x.br; ++r) {
      const char *pa, *pb; br_ptrs(d, x, r, 1, 1, pa, pb);
      for (int q0 = 0; q0 < kq_all; q0 += kcq) {
        const int qn = (kq_all - q0 < kcq) ? (kq_all - q0) : kcq;
        group_sync(wpt, group);                        // previous chunk fully consumed
        for (int e = gtid; e < qn * m; e += gsize) {   // A words: rows of m contiguous words
          const int q = e / m, i = e - q * m;
          sa[q * ms + i] = reinterpret_cast<const unsigned int*>(pa + ((long long)(q0 + q) * lda) * 4)[i];
        }
        for (int e = gtid; e < n * qn; e += gsize) {   // B words: rows of qn contiguous words
          const int jn = e / qn, q = e - jn * qn;
          sb[q * ns + jn] = reinterpret_cast<const unsigned int*>(pb + (long long)jn * ldb + (long long)q0 * 4)[q];
        }
        group_sync(wpt, group);
        if (n0 < n) for (int q = 0; q < qn; ++q) {
          unsigned int av[TM], bv[TN];
          if (TM == 4) {   // operands beyond m / n are padding words of the row (never stored to C)
            const uint4 a4 = *reinterpret_cast<const uint4*>(sa + q * ms + m0 + lm * 4);
            const uint4 b0 = *reinterpret_cast<const uint4*>(sb + q * ns + n0 + ln * 8), b1 = *reinterpret_cast<const uint4*>(sb + q * ns + n0 + ln * 8 + 4);
            av[0] = a4.x; av[1] = a4.y; av[2] = a4.z; av[3 % TM] = a4.w;
            bv[0] = b0.x; bv[1] = b0.y; bv[2 % TN] = b0.z; bv[3 % TN] = b0.w; bv[4 % TN] = b1.x; bv[5 % TN] = b1.y; bv[6 % TN] = b1.z; bv[7 % TN] = b1.w;
          } else {
#pragma unroll
            for (int i = 0; i < TM; ++i) { const int mi = m0 + lm * TM + i; av[i] = (mi < m) ? sa[q * ms + mi] : 0u; }
#pragma unroll
            for (int j = 0; j < TN; ++j) { const int nj = n0 + ln * TN + j; bv[j] = (nj < n) ? sb[q * ns + nj] : 0u; }
          }
#pragma unroll
          for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = dp4a_x<UA, UB>(av[i], bv[j], acc[i][j]);
        }
      }
    }
    if (live && TM == 4 && (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(x.c) & 15) == 0 && m0 + lm * 4 + 3 < m) {
      // four consecutive rows per lane: 16-byte accesses, 8 lanes cover 128 contiguous bytes of a C column
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int nj = n0 + ln * TN + j;
        if (nj < n) {
          const long long ci = (long long)nj * ldc + m0 + lm * 4;
          if (!to_f32) {
            uint4 v = make_uint4(acc[0][j], acc[1 % TM][j], acc[2 % TM][j], acc[3 % TM][j]);
            if (!beta0) { const uint4 o = *reinterpret_cast<const uint4*>(reinterpret_cast<const int*>(x.c) + ci); v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
            *reinterpret_cast<uint4*>(reinterpret_cast<int*>(x.c) + ci) = v;
          } else {
            float4 f = make_float4(__fmul_rn((float)(int)acc[0][j], x.scf), __fmul_rn((float)(int)acc[1 % TM][j], x.scf),
                                   __fmul_rn((float)(int)acc[2 % TM][j], x.scf), __fmul_rn((float)(int)acc[3 % TM][j], x.scf));
            if (!beta0) { const float4 o = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x.c) + ci);
                          f.x = __fadd_rn(f.x, o.x); f.y = __fadd_rn(f.y, o.y); f.z = __fadd_rn(f.z, o.z); f.w = __fadd_rn(f.w, o.w); }
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(x.c) + ci) = f;
          }
        }
      }
    } else if (live) {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int nj = n0 + ln * TN + j;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const int mi = m0 + lm * TM + i;
          if (mi < m && nj < n) {
            const long long ci = (long long)nj * ldc + mi;
            if (!to_f32) {
              unsigned int v = acc[i][j];
              if (!beta0) v += (unsigned int)reinterpret_cast<const int*>(x.c)[ci];
              reinterpret_cast<int*>(x.c)[ci] = (int)v;
            } else {                                               // reference :1579-1585
              float f = __fmul_rn((float)(int)acc[i][j], x.scf);
              if (!beta0) f = __fadd_rn(f, reinterpret_cast<const float*>(x.c)[ci]);
              reinterpret_cast<float*>(x.c)[ci] = f;
            }
          }
        }
      }
    }
  }
}

// host side: is the dp4a kernel applicable? (VNNI4 A, whole words everywhere, strided or single-tile launch)
bool i8_fast_ok(const xb_gemm_launch& L, int path) {
  const xb_gemm_desc& d = L.d;
  if (path != P_I8_I32 && path != P_I8_F32) return false;
  if (path == P_I8_I32 && (d.flags & LIBXSMM_GEMM_FLAG_VNNI_A) == 0) return false;    // flat A: bytes of 4 different rows per word
  if ((d.k & 3) != 0 || (d.ldb & 3) != 0 || d.m > 1024 || d.n > 1024) return false;
  if (L.recs != nullptr || d.br_type == 1 || d.br_type == 2) return false;            // per-tile pointers are not checkable on the host
  const bool single = (L.a == nullptr && L.c == nullptr);
  const uintptr_t a = (uintptr_t)(single ? L.one.a : L.a), b = (uintptr_t)(single ? L.one.b : L.b), c = (uintptr_t)(single ? L.one.c : L.c);
  if (((a | b | c) & 3) != 0) return false;
  if (!single && (((L.tile_stride_a | L.tile_stride_b | L.tile_stride_c) & 3) != 0)) return false;
  if (d.br_type == 3 && (((d.br_stride_a | d.br_stride_b) & 3) != 0)) return false;
  return true;
}

template <int TM, int TN>
cudaError_t launch_i8(const xb_gemm_launch& L, int path, int wpt, int words_per_group, size_t smem, unsigned int grid, cudaStream_t st) {
  const bool ua = (L.d.ta == LIBXSMM_DATATYPE_U8), ub = (L.d.tb == LIBXSMM_DATATYPE_U8);
  const int to_f32 = (path == P_I8_F32);
#define XB_I8_CASE(A, B) do { \
    static unsigned long long attr_set = 0ull; \
    if (xb_rt_first_use_on_device(&attr_set)) cudaFuncSetAttribute(gemm_i8_kernel<TM, TN, A, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
    gemm_i8_kernel<TM, TN, A, B><<<grid, I8_WARPS * 32, smem, st>>>(L, to_f32, wpt, words_per_group); } while (0)
  if (ua && ub) XB_I8_CASE(true, true); else if (ua) XB_I8_CASE(true, false); else if (ub) XB_I8_CASE(false, true); else XB_I8_CASE(false, false);
#undef XB_I8_CASE
  return cudaGetLastError();
}
}  // namespace

extern "C" int xb_gemm_simt_supported(const xb_gemm_desc* d) {
  const int path = xb_path_of(*d);
  if (path == P_NONE) return 0;
  if (d->fuse_colbias != 0 || d->cp_op != 0) {   // fused form: float C only (the reference's f32 image of C)
    return path == P_F32 || path == P_I8_F32 || path == P_F16_F16 || path == P_F16_F32 || path == P_BF16_F32 || path == P_BF16_BF16;
  }
  return 1;
}

extern "C" int xb_gemm_simt_launch(const xb_gemm_launch* L) {
  const int path = xb_path_of(L->d);
  if (path == P_NONE) return 1;
  if (L->count <= 0) return 0;
  if (path == P_BITMAP) {
    if (L->count != 1 || L->recs != nullptr || L->one.a_q == nullptr) { xb_rt_note_error(1, "bitmap-compressed A: single calls only"); return 1; }
    unsigned int* prefix = (unsigned int*)xb_rt_scratch((size_t)(((long long)L->d.m * L->d.k + 31) / 32) * 4 + 16);
    if (prefix == nullptr) return 2;
    gemm_bitmap_kernel<<<1, 1024, 0, (cudaStream_t)xb_rt_stream()>>>(*L, prefix);
    xb_rt_count_launch();
    const cudaError_t be = cudaGetLastError();
    if (be != cudaSuccess) { xb_rt_note_error((int)be, "gemm_bitmap"); return (int)be; }
    return 0;
  }
  if (path == P_I4_I32 && L->one.a_q == nullptr && L->recs == nullptr) { xb_rt_note_error(1, "int4 A: zero points missing (a.quaternary)"); return 1; }
  if (L->d.fuse_colbias != 0 || L->d.cp_op != 0) {
    const long long fgrid = L->count < (1 << 20) ? L->count : (1 << 20);
    gemm_simt_fused_kernel<<<(unsigned int)fgrid, 256, 0, (cudaStream_t)xb_rt_stream()>>>(*L, path);
    xb_rt_count_launch();
    const cudaError_t fe = cudaGetLastError();
    if (fe != cudaSuccess) { xb_rt_note_error((int)fe, "gemm_simt_fused"); return (int)fe; }
    return 0;
  }
  if (i8_fast_ok(*L, path) && getenv("LIBXSMM_B200_I8_EXACT_ORDER") == nullptr) {
    const int small = (L->d.m < L->d.n) ? L->d.m : L->d.n;
    const int tm = (small >= 32) ? 4 : ((small >= 16) ? 2 : 1), tn = 2 * tm;           // lane block; a warp covers 8*tm x 4*tn of C
    const int passes = ((L->d.m + 8 * tm - 1) / (8 * tm)) * ((L->d.n + 4 * tn - 1) / (4 * tn));
    int wpt = 1; while (wpt < passes) wpt *= 2;                                        // warps per tile: 1, 2, 4, 8 or 16
    const int kcq = (L->d.k / 4 < I8_KC / 4) ? L->d.k / 4 : I8_KC / 4;
    // panels [kcq][ms] + [kcq][ns] (+ slack: a 4x8 lane block may read up to 31 padding words past the last row)
    const int words = kcq * (((((L->d.m + 3) >> 2) | 1) << 2) + ((((L->d.n + 3) >> 2) | 1) << 2)) + 64;
    const int groups = I8_WARPS / (wpt > I8_WARPS ? I8_WARPS : wpt);
    const size_t smem = (size_t)words * 4 * groups;
    if (wpt <= I8_WARPS && smem <= 200 * 1024) {
      static int sms = 0;
      if (sms == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
      const long long ctas_needed = (L->count + groups - 1) / groups;
      long long per_sm = (long long)((220 * 1024) / (smem + 1024)); if (per_sm < 1) per_sm = 1; if (per_sm > 2) per_sm = 2;   // 64 registers x 512 threads: two CTAs per SM
      long long grid = (long long)sms * per_sm;
      if (grid > ctas_needed) grid = ctas_needed;
      cudaStream_t st = (cudaStream_t)xb_rt_stream();
      cudaError_t e;
      if (tm == 4) e = launch_i8<4, 8>(*L, path, wpt, words, smem, (unsigned int)grid, st);
      else if (tm == 2) e = launch_i8<2, 4>(*L, path, wpt, words, smem, (unsigned int)grid, st);
      else e = launch_i8<1, 2>(*L, path, wpt, words, smem, (unsigned int)grid, st);
      xb_rt_count_launch();
      if (e != cudaSuccess) { xb_rt_note_error((int)e, "gemm_i8"); return (int)e; }
      return 0;
    }
  }
  const long long grid = L->count < (1 << 20) ? L->count : (1 << 20);
  gemm_simt_kernel<<<(unsigned int)grid, 256, 0, (cudaStream_t)xb_rt_stream()>>>(*L, path);
  xb_rt_count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { xb_rt_note_error((int)e, "gemm_simt"); return (int)e; }
  return 0;
}
