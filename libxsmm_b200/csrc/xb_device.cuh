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

// bf8 (E5M2) is the upper byte of an IEEE half: convert to half, round the lower byte away to nearest even, keep
// Inf, quiet NaN (reference src/libxsmm_math.c:731-746)
XB_HD uint8_t xb_f32_to_bf8(float f) {
  unsigned short h = xb_f32_to_f16(f);
  if ((h & 0x7c00) == 0x7c00) { if ((h & 0x03ff) != 0) h |= 0x0200; }
  else h = (unsigned short)(h + 0x007f + ((h >> 8) & 1));
  return (uint8_t)(h >> 8);
}
// hf8 (E4M3, bias 7, no infinities: 0x7f is NaN, max normal 448): via half with nearest-even on the dropped 7 mantissa
// bits, subnormals for exponents below the bias, overflow -> NaN (reference src/libxsmm_math.c:749-822)
XB_HD uint8_t xb_f32_to_hf8(float f) {
  unsigned short h = xb_f32_to_f16(f);
  const unsigned short sign = (unsigned short)((h & 0x8000) >> 8);
  const unsigned int e16 = (h & 0x7c00u) >> 10, m16 = h & 0x03ffu;
  unsigned int e, m;
  if (e16 == 0x1f) { e = 0xf; m = 0x7; }
  else if (e16 > 23 || (e16 == 23 && m16 > 0x0340)) { e = 0xf; m = 0x7; }      // beyond 448 (+ half an ulp)
  else if (e16 < 5) { e = 0; m = 0; }                                              // below half the smallest subnormal
  else if (e16 <= 8) {                                                             // subnormal result
    m = (m16 | 0x0400u) >> (9 - e16);
    m |= ((m16 & 0x007fu) + 0x007fu) >> 7;                                         // sticky bit of what the shift dropped
    m = (m + 0x003fu + ((m >> 7) & 1u)) >> 7;
    e = 0;
  } else {
    h = (unsigned short)(h + 0x003f + ((m16 >> 7) & 1u));
    e = ((h & 0x7c00u) >> 10) - 8; m = (h & 0x03ffu) >> 7;
  }
  return (uint8_t)(sign | (e << 3) | m);
}
XB_HD float xb_hf8_to_f32(uint8_t in) {
  const unsigned int sign = ((unsigned int)in & 0x80u) << 24, e = ((unsigned int)in & 0x78u) >> 3;
  unsigned int m = (unsigned int)in & 0x07u, e32 = e + 120;
  if (e == 0 && m != 0) {                          // subnormal: renormalise
    const unsigned int lz = (m > 3) ? 0 : ((m > 1) ? 1 : 2);
    e32 -= lz; m = (m << (lz + 1)) & 0x07u;
  } else if (e == 0) e32 = 0;
  else if (e == 0xf && m == 0x7) { e32 = 0xff; m = 0x4; }
  return xb_bits_f32(sign | (e32 << 23) | (m << 20));
}
XB_HD float xb_bf8_to_f32(uint8_t in) { return xb_f16_to_f32((uint16_t)((uint16_t)in << 8)); }   // libxsmm_convert_bf8_to_f32 (:546-551)

#endif  // XB_DEVICE_CUH
