"""Seeded input generators shared by the parity tests and bench.py.

Value distributions follow the reference drivers: multiples of 0.1 in [-0.5, 0.5] for floating point
(samples/xgemm_sparse/spmm_kernel.c:498-527), that times 40 for 8-bit integers (:594-599)."""
import numpy as np

F64, F32, BF16, F16, I32, I16, I8, U8 = 0, 1, 2, 3, 8, 10, 12, 13
BF8, HF8, MXBF8, MXFP4X2, NVFP4X2 = 4, 5, 14, 20, 21
NP_OF = {F64: np.float64, F32: np.float32, BF16: np.uint16, F16: np.uint16, I32: np.int32, I16: np.int16, I8: np.int8, U8: np.uint8, 4: np.uint8, 5: np.uint8}
TS = {F64: 8, F32: 4, BF16: 2, F16: 2, I32: 4, I16: 2, I8: 1, U8: 1, 4: 1, 5: 1}


def f32_to_bf16_bits(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = u + 0x7FFF + ((u >> 16) & 1)
    return (u >> 16).astype(np.uint16)


def bf16_bits_to_f32(h):
    return (np.asarray(h, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def values(rng, n, dtype):
    """n elements of `dtype` (libxsmm datatype id) drawn like the reference drivers do."""
    tenths = rng.integers(-5, 6, size=n)
    if dtype == F64:
        return (tenths / 10.0).astype(np.float64)
    if dtype == F32:
        return (tenths / 10.0).astype(np.float32)
    if dtype == BF16:
        return f32_to_bf16_bits((tenths / 10.0).astype(np.float32))
    if dtype == F16:
        return (tenths / 10.0).astype(np.float16).view(np.uint16)
    if dtype == I8:
        return (tenths * 4).astype(np.int8)          # 0.1 * 40
    if dtype == U8:
        return (np.abs(tenths) * 4).astype(np.uint8)
    if dtype == I16:
        return (tenths * 40).astype(np.int16)
    if dtype == I32:
        return rng.integers(-1000, 1000, size=n).astype(np.int32)
    if dtype == BF8:       # E5M2 = upper byte of an f16: the tenths, truncated
        return ((tenths / 10.0).astype(np.float16).view(np.uint16) >> 8).astype(np.uint8)
    if dtype == HF8:       # E4M3: sign, exponent 4..9 (1/8 .. 7.5), any mantissa; no NaN code
        return ((rng.integers(0, 2, size=n) << 7) | (rng.integers(4, 10, size=n) << 3) | rng.integers(0, 8, size=n)).astype(np.uint8)
    raise ValueError(dtype)


def to_f64(arr, dtype):
    if dtype == BF16:
        return bf16_bits_to_f32(arr).astype(np.float64)
    if dtype == F16:
        return arr.view(np.float16).astype(np.float64)
    if dtype == BF8:
        return (arr.astype(np.uint16) << 8).view(np.float16).astype(np.float64)
    if dtype == HF8:
        b = arr.astype(np.int64); e, m = (b >> 3) & 15, b & 7
        v = np.where(e == 0, m * 2.0 ** -9, (8 + m) * np.exp2((e - 10).astype(np.float64)))
        v = np.where((e == 15) & (m == 7), np.nan, v)
        return np.where(b & 0x80, -v, v)
    return arr.astype(np.float64)


def normf_rel(ref, tst):
    """relative Frobenius-norm error, the acceptance norm of samples/xgemm/gemm_kernel.c:5312-5414."""
    ref = np.asarray(ref, dtype=np.float64).ravel()
    tst = np.asarray(tst, dtype=np.float64).ravel()
    den = np.linalg.norm(ref)
    return float(np.linalg.norm(ref - tst) / den) if den > 0 else float(np.linalg.norm(tst))
