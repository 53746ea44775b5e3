// libxsmm_b200 -- device/host helpers shared by all kernels: element sizes and the low-precision
// conversions whose rounding rules the kernels must match bit for bit.
// Rounding rules follow the reference's software conversions (src/libxsmm_math.c):
//   f32 -> bf16 : :684-703  flush f32 denormals to signed zero, round to nearest even, quiet NaNs
//   bf16 -> f32 : GEMM kernels widen by a plain 16-bit shift (generator_gemm_reference_impl.c:2141-2165)
//   f32 -> f16  : :824-900  IEEE round-to-nearest-even incl. subnormal results, NaN keeps top payload
//   f16 -> f32  : :600-640  exact widening incl. subnormals
#ifndef XB_DEVICE_CUH
#define XB_DEVICE_CUH

#include <stdint.h>
#include "../../include/libxsmm_typedefs.h"

#if defined(__CUDACC__)
# define XB_HD __host__ __device__ __forceinline__
#else
# define XB_HD static inline
#endif

XB_HD int xb_dev_typesize(int t) {
  switch (t) {
    case LIBXSMM_DATATYPE_F64: case LIBXSMM_DATATYPE_I64: case LIBXSMM_DATATYPE_U64: return 8;
    case LIBXSMM_DATATYPE_F32: case LIBXSMM_DATATYPE_I32: case LIBXSMM_DATATYPE_U32: case LIBXSMM_DATATYPE_BF32: return 4;
    case LIBXSMM_DATATYPE_BF16: case LIBXSMM_DATATYPE_F16: case LIBXSMM_DATATYPE_I16: case LIBXSMM_DATATYPE_U16: return 2;
    case LIBXSMM_DATATYPE_IMPLICIT: case LIBXSMM_DATATYPE_UNSUPPORTED: return 0;
    default: return 1;
  }
}

XB_HD uint32_t xb_f32_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
XB_HD float xb_bits_f32(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

XB_HD float xb_bf16_to_f32(uint16_t h) { return xb_bits_f32(((uint32_t)h) << 16); }

XB_HD uint16_t xb_f32_to_bf16_rne(float f) {
  uint32_t u = xb_f32_bits(f);
  const uint32_t expo = u & 0x7f800000u;
  if (expo == 0u) u &= 0x80000000u;                       // denormal in -> signed zero
  if (expo == 0x7f800000u) {                              // inf stays, NaN gets its quiet bit
    if ((u & 0x007fffffu) != 0u) u |= 0x00400000u;
  } else {
    u += 0x7fffu + ((u >> 16) & 1u);                      // nearest, ties to even
  }
  return (uint16_t)(u >> 16);
}

XB_HD float xb_f16_to_f32(uint16_t h) {
  const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  uint32_t expo = ((uint32_t)h >> 10) & 0x1fu;
  uint32_t mant = (uint32_t)h & 0x3ffu;
  uint32_t out;
  if (expo == 0x1fu) {                                    // inf / NaN (NaN quieted)
    if (mant != 0u) mant |= 0x200u;
    out = 0x7f800000u | (mant << 13);
  } else if (expo == 0u) {
    if (mant == 0u) out = 0u;
    else {                                                // subnormal: renormalise
      int shift = 0;
      while ((mant & 0x400u) == 0u) { mant <<= 1; ++shift; }
      mant &= 0x3ffu;
      out = ((uint32_t)(113 - shift) << 23) | (mant << 13);
    }
  } else {
    out = ((expo + 112u) << 23) | (mant << 13);
  }
  return xb_bits_f32(out | sign);
}

XB_HD uint16_t xb_f32_to_f16(float f) {
  uint32_t u = xb_f32_bits(f);
  const uint32_t sign = (u & 0x80000000u) >> 16;
  const uint32_t e32 = (u >> 23) & 0xffu;
  const uint32_t m32 = u & 0x007fffffu;
  uint32_t e, m;
  if (e32 == 0xffu) {                                     // inf / NaN
    e = 0x1fu; m = (m32 == 0u) ? 0u : ((m32 >> 13) | 0x200u);
  } else if (e32 > 142u) {                                // |x| >= 2^16 -> inf
    e = 0x1fu; m = 0u;
  } else if (e32 < 102u) {                                // below half the smallest subnormal (and DAZ)
    e = 0u; m = 0u;
  } else if (e32 <= 112u) {                               // subnormal result, nearest-even with sticky
    uint32_t mm = (m32 | 0x00800000u) >> (113u - e32);
    mm |= (((m32 & 0x1fffu) + 0x1fffu) >> 13);
    mm += 0xfffu + ((mm >> 13) & 1u);
    m = mm >> 13; e = 0u;
    return (uint16_t)(sign | m);                          // a carry into bit 10 yields the smallest normal
  } else {
    const uint32_t r = (u & 0x7fffffffu) + 0xfffu + ((m32 >> 13) & 1u);
    e = ((r >> 23) & 0xffu) - 112u; m = (r & 0x007fffffu) >> 13;
    return (uint16_t)(sign | (e << 10) | m);
  }
  return (uint16_t)(sign | (e << 10) | m);
}

#endif  // XB_DEVICE_CUH
