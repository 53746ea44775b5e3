// libxsmm_b200 -- accumulator write-back shared by the tcgen05 GEMM kernels (gemm_tc.cu, gemm_ts.cu).
// A thread owns one row of the tile and holds 32 consecutive columns of it (one tcgen05.ld 32x32b.x32). Lanes are
// consecutive rows, so every store instruction of a warp writes one contiguous run of a column of the column-major C.
// The output type and beta are template parameters and the address is a running pointer: the per-element work is the
// conversion, one 64-bit add and the store -- nothing is re-derived from kernel parameters inside the loop.
// Semantics per type follow libxsmm_ref_matmul (src/generator_gemm_reference_impl.c): f32 accumulate, one rounding at
// the end (:2367-2419 bf16, :2025-2126 f16), beta = 1 adds the old C in f32 (f16 inputs with f32 C round the old value
// through f16 first, :2118-2123), int32 wraps (:1452-1555), int8 -> f32 scales by scf (:1556-1683).
#ifndef XB_EPILOGUE_CUH
#define XB_EPILOGUE_CUH
#include "xb_device.cuh"

enum { XB_EP_F32 = 0, XB_EP_F32_FROM_F16 = 1, XB_EP_BF16 = 2, XB_EP_F16 = 3, XB_EP_I32 = 4, XB_EP_I32_TO_F32 = 5 };

template <int MODE, bool BETA0>
__device__ __forceinline__ void xb_ep_store_one(uint32_t raw, char* p, float scf) {
  if (MODE == XB_EP_F32 || MODE == XB_EP_F32_FROM_F16) {
    float acc = __uint_as_float(raw);
    float* d = reinterpret_cast<float*>(p);
    if (!BETA0) { float old = *d; if (MODE == XB_EP_F32_FROM_F16) old = xb_f16_to_f32(xb_f32_to_f16(old)); acc += old; }
    *d = acc;
  } else if (MODE == XB_EP_BF16) {
    float acc = __uint_as_float(raw);
    unsigned short* d = reinterpret_cast<unsigned short*>(p);
    if (!BETA0) acc += xb_bf16_to_f32(*d);
    *d = xb_f32_to_bf16_rne(acc);
  } else if (MODE == XB_EP_F16) {
    float acc = __uint_as_float(raw);
    unsigned short* d = reinterpret_cast<unsigned short*>(p);
    if (!BETA0) acc += xb_f16_to_f32(*d);
    *d = xb_f32_to_f16(acc);
  } else if (MODE == XB_EP_I32) {
    unsigned int* d = reinterpret_cast<unsigned int*>(p);
    *d = BETA0 ? raw : raw + *d;
  } else {
    float* d = reinterpret_cast<float*>(p);
    float f = __fmul_rn((float)(int)raw, scf);
    if (!BETA0) f = __fadd_rn(f, *d);
    *d = f;
  }
}

// bf16 output, full chunk: the software RNE of xb_f32_to_bf16_rne costs ~16 instructions per value; the hardware conversion
// (cvt.rn.bf16x2.f32, two values per instruction) gives the same bits for every zero, normal and infinite input. Where
// libxsmm_convert_f32_to_bf16_rne differs: f32 denormals are flushed to signed zero first -- a multiply by one in .ftz mode
// does exactly that -- and NaNs keep their upper payload bits; but an accumulator NaN on this machine is always the canonical
// 0x7fffffff (tensor core and FADD both canonicalise), for which both conversions give 0x7fff.
template <bool BETA0>
__device__ __forceinline__ void xb_ep_store_chunk_bf16_full(const uint32_t (&v)[32], char* p, long long ldc_bytes) {
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    float a0 = __uint_as_float(v[j]), a1 = __uint_as_float(v[j + 1]);
    if (!BETA0) {
      a0 += xb_bf16_to_f32(*reinterpret_cast<const unsigned short*>(p));
      a1 += xb_bf16_to_f32(*reinterpret_cast<const unsigned short*>(p + ldc_bytes));
    }
    float f0, f1; uint32_t two;
    asm("mul.ftz.f32 %0, %1, 0f3F800000;" : "=f"(f0) : "f"(a0));
    asm("mul.ftz.f32 %0, %1, 0f3F800000;" : "=f"(f1) : "f"(a1));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(two) : "f"(f1), "f"(f0));                 // upper half <- f1, lower half <- f0
    *reinterpret_cast<unsigned short*>(p) = (unsigned short)two; p += ldc_bytes;
    *reinterpret_cast<unsigned short*>(p) = (unsigned short)(two >> 16); p += ldc_bytes;
  }
}

// p: address of (this thread's row, first column of the chunk); ncols: valid columns of the chunk (1..32)
template <int MODE, bool BETA0>
__device__ __forceinline__ void xb_ep_store_chunk_t(const uint32_t (&v)[32], char* p, long long ldc_bytes, int ncols, float scf) {
  if (MODE == XB_EP_BF16 && ncols >= 32) { xb_ep_store_chunk_bf16_full<BETA0>(v, p, ldc_bytes); return; }
  if (ncols >= 32) {
#pragma unroll
    for (int j = 0; j < 32; ++j) { xb_ep_store_one<MODE, BETA0>(v[j], p, scf); p += ldc_bytes; }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) { if (j < ncols) xb_ep_store_one<MODE, BETA0>(v[j], p, scf); p += ldc_bytes; }
  }
}

__device__ __forceinline__ void xb_ep_store_chunk(int mode, int beta0, const uint32_t (&v)[32], char* p, long long ldc_bytes, int ncols, float scf) {
  switch (mode * 2 + (beta0 ? 1 : 0)) {
    case XB_EP_F32 * 2 + 0:          xb_ep_store_chunk_t<XB_EP_F32, false>(v, p, ldc_bytes, ncols, scf); break;
    case XB_EP_F32 * 2 + 1:          xb_ep_store_chunk_t<XB_EP_F32, true>(v, p, ldc_bytes, ncols, scf); break;
    case XB_EP_F32_FROM_F16 * 2 + 0: xb_ep_store_chunk_t<XB_EP_F32_FROM_F16, false>(v, p, ldc_bytes, ncols, scf); break;
    case XB_EP_F32_FROM_F16 * 2 + 1: xb_ep_store_chunk_t<XB_EP_F32, true>(v, p, ldc_bytes, ncols, scf); break;
    case XB_EP_BF16 * 2 + 0:         xb_ep_store_chunk_t<XB_EP_BF16, false>(v, p, ldc_bytes, ncols, scf); break;
    case XB_EP_BF16 * 2 + 1:         xb_ep_store_chunk_t<XB_EP_BF16, true>(v, p, ldc_bytes, ncols, scf); break;
    case XB_EP_F16 * 2 + 0:          xb_ep_store_chunk_t<XB_EP_F16, false>(v, p, ldc_bytes, ncols, scf); break;
    case XB_EP_F16 * 2 + 1:          xb_ep_store_chunk_t<XB_EP_F16, true>(v, p, ldc_bytes, ncols, scf); break;
    case XB_EP_I32 * 2 + 0:          xb_ep_store_chunk_t<XB_EP_I32, false>(v, p, ldc_bytes, ncols, scf); break;
    case XB_EP_I32 * 2 + 1:          xb_ep_store_chunk_t<XB_EP_I32, true>(v, p, ldc_bytes, ncols, scf); break;
    case XB_EP_I32_TO_F32 * 2 + 0:   xb_ep_store_chunk_t<XB_EP_I32_TO_F32, false>(v, p, ldc_bytes, ncols, scf); break;
    default:                         xb_ep_store_chunk_t<XB_EP_I32_TO_F32, true>(v, p, ldc_bytes, ncols, scf); break;
  }
}

// host side: epilogue mode and C element size of a GEMM descriptor's (a type, c type) pair
static inline int xb_ep_mode(int a_type, int c_type, int* esz) {
  const int a8 = (a_type == LIBXSMM_DATATYPE_I8 || a_type == LIBXSMM_DATATYPE_U8);
  if (a8) { *esz = 4; return (c_type == LIBXSMM_DATATYPE_I32) ? XB_EP_I32 : XB_EP_I32_TO_F32; }
  if (c_type == LIBXSMM_DATATYPE_F32) { *esz = 4; return (a_type == LIBXSMM_DATATYPE_F16) ? XB_EP_F32_FROM_F16 : XB_EP_F32; }
  *esz = 2;
  return (c_type == LIBXSMM_DATATYPE_BF16) ? XB_EP_BF16 : XB_EP_F16;
}
#endif  // XB_EPILOGUE_CUH
