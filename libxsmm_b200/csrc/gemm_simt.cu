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
#include "xb_internal.h"
#include "xb_device.cuh"

namespace {

enum { P_F64 = 0, P_F32, P_I16, P_I8_I32, P_I8_F32, P_F16_F16, P_F16_F32, P_BF16_F32, P_BF16_BF16, P_NONE };

__host__ __device__ inline int xb_path_of(const xb_gemm_desc& d) {
  const int a = d.ta, b = d.tb, c = d.tc, comp = d.tcomp;
  const bool a8 = (a == LIBXSMM_DATATYPE_I8 || a == LIBXSMM_DATATYPE_U8);
  const bool b8 = (b == LIBXSMM_DATATYPE_I8 || b == LIBXSMM_DATATYPE_U8);
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
  return P_NONE;
}

struct TileCtx {
  const char* a0; const char* b0; char* c;       // tile bases
  const void* const* a_addr; const void* const* b_addr;   // address mode
  const long long* a_offs; const long long* b_offs;       // offset mode
  unsigned long long br;
  float scf;
};

__device__ inline void resolve_tile(const xb_gemm_launch& L, long long t, TileCtx& x) {
  xb_gemm_rec r;
  if (L.recs != nullptr) r = L.recs[t];
  else if (L.a == nullptr && L.c == nullptr) r = L.one;
  else {
    r.a = (const char*)L.a + t * L.tile_stride_a; r.b = (const char*)L.b + t * L.tile_stride_b;
    r.c = (char*)L.c + t * L.tile_stride_c; r.a_aux = nullptr; r.b_aux = nullptr; r.br = L.br; r.scf = L.one.scf;
    if (L.d.br_type == 2) { r.a_aux = L.one.a_aux; r.b_aux = L.one.b_aux; }
  }
  x.a0 = (const char*)r.a; x.b0 = (const char*)r.b; x.c = (char*)r.c;
  x.a_addr = (const void* const*)r.a; x.b_addr = (const void* const*)r.b;
  x.a_offs = (const long long*)r.a_aux; x.b_offs = (const long long*)r.b_aux;
  x.br = (L.d.br_type == 0) ? 1ull : r.br; x.scf = r.scf;
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
        default: break;
      }
    }
  }
}

}  // namespace

extern "C" int xb_gemm_simt_supported(const xb_gemm_desc* d) { return xb_path_of(*d) != P_NONE; }

extern "C" int xb_gemm_simt_launch(const xb_gemm_launch* L) {
  const int path = xb_path_of(L->d);
  if (path == P_NONE) return 1;
  if (L->count <= 0) return 0;
  const long long grid = L->count < (1 << 20) ? L->count : (1 << 20);
  gemm_simt_kernel<<<(unsigned int)grid, 256, 0, (cudaStream_t)xb_rt_stream()>>>(*L, path);
  xb_rt_count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { xb_rt_note_error((int)e, "gemm_simt"); return (int)e; }
  return 0;
}
