"""TEST INFRASTRUCTURE: ctypes access to the two CPU checkers.

  oracle  -> oracle/liboracle.so        plain-C restatement of the reference algorithms (oracle/oracle.c)
  ref     -> oracle/_ref/libxsmm_ref.so the unmodified reference (header-only build, oracle/ref_shim.c)

Both expose the same call shapes (oracle_* / ref_*). Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libxsmm_ref.so")

_I, _U, _P, _LL, _ULL, _F = C.c_int, C.c_uint, C.c_void_p, C.c_longlong, C.c_ulonglong, C.c_float


def _build_if_needed():
    srcs = [os.path.join(ROOT, "oracle", f) for f in ("oracle.c", "oracle_meltw.c")]
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", ROOT, "oracle"], stdout=subprocess.DEVNULL)
    if not os.path.exists(REF_SO) and os.path.isdir("/root/reference/include"):
        subprocess.check_call(["make", "-C", ROOT, "ref"], stdout=subprocess.DEVNULL)


def _bind(lib, prefix):
    ns = {}

    def sig(name, restype, argtypes):
        fn = getattr(lib, prefix + name)
        fn.restype, fn.argtypes = restype, argtypes
        ns[name] = fn

    sig("gemm", _I, [_P, _P, _U, _I, _LL, _LL, _ULL, _P, _P, _P, _P, _P, _F, _I])
    sig("fsspmdm", _I, [_I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P])
    sig("bcsc", _I, [_P, _P, _U, _P, _P, _P, _P, _P])
    sig("packed_sp", _I, [_I, _I, _P, _U, _I, _P, _P, _P, _P, _P, _P])
    sig("packed_dense", _I, [_I, _I, _P, _U, _I, _P, _P, _P])
    return ns


_build_if_needed()
oracle_lib = C.CDLL(ORACLE_SO)
oracle = _bind(oracle_lib, "oracle_")
for _n, _r, _a in (("f32_to_bf16", C.c_ushort, [_F]), ("f32_to_f16", C.c_ushort, [_F]), ("f16_to_f32", _F, [C.c_ushort]),
                   ("bf16_to_f32", _F, [C.c_ushort])):
    _fn = getattr(oracle_lib, "oracle_" + _n)
    _fn.restype, _fn.argtypes = _r, _a
    oracle[_n] = _fn
oracle_lib.oracle_meltw.restype, oracle_lib.oracle_meltw.argtypes = _I, [_P, _P, _I]
oracle["meltw"] = oracle_lib.oracle_meltw          # 0 = computed, 2 = op not restated (the reference stays the only checker)
oracle_lib.oracle_gemm_i4.restype = _I
oracle_lib.oracle_gemm_i4.argtypes = [_P, _U, _I, _LL, _LL, _ULL, _P, _P, _P, _P]
oracle["gemm_i4"] = oracle_lib.oracle_gemm_i4
oracle_lib.oracle_gemm_bitmap.restype = _I
oracle_lib.oracle_gemm_bitmap.argtypes = [_P, _P, _U, _P, _P, _P, _P]
oracle["gemm_bitmap"] = oracle_lib.oracle_gemm_bitmap
oracle_lib.oracle_gemm_ext.restype = _I
oracle_lib.oracle_gemm_ext.argtypes = [_P, _P, _U, _I, _LL, _LL, _ULL, _P, _P, _P, _P, _P, _F, _P, _P, _P]
oracle["gemm_ext"] = oracle_lib.oracle_gemm_ext
_fn = oracle_lib.oracle_gemm_batch
_fn.restype, _fn.argtypes = _I, [_P, _P, _U, _I, _LL, _LL, _ULL, _P, _P, _P, _LL, _LL, _LL, _LL]
oracle["gemm_batch"] = _fn

ref = None
ref_lib = None
if os.path.exists(REF_SO):
    ref_lib = C.CDLL(REF_SO)
    ref = _bind(ref_lib, "ref_")
    for _n, _r, _a in (("f32_to_bf16", C.c_ushort, [_F]), ("f32_to_f16", C.c_ushort, [_F]), ("f16_to_f32", _F, [C.c_ushort]),
                       ("bf16_to_f32", _F, [C.c_ushort])):
        _fn = getattr(ref_lib, "ref_" + _n)
        _fn.restype, _fn.argtypes = _r, _a
        ref[_n] = _fn
    ref_lib.ref_meltw.restype, ref_lib.ref_meltw.argtypes = _I, [_P, _P, _I]
    ref["meltw"] = ref_lib.ref_meltw
    ref_lib.ref_meqn.restype = _I
    ref_lib.ref_meqn.argtypes = [_P, _I, _P, _P, _I, _P]
    ref["meqn"] = ref_lib.ref_meqn
    ref_lib.ref_gemm_aux.restype = _I
    ref_lib.ref_gemm_aux.argtypes = [_P, _P, _U, _I, _LL, _LL, _ULL, _P, _P, _P, _I, _P]
    ref["gemm_aux"] = ref_lib.ref_gemm_aux
    ref_lib.ref_gemm_ext.restype = _I
    ref_lib.ref_gemm_ext.argtypes = [_P, _P, _U, _I, _LL, _LL, _ULL, _P, _P, _P, _P, _P, _F, _P, _P, _P, _I]
    ref["gemm_ext"] = lambda *a: ref_lib.ref_gemm_ext(*a, 0)
    ref_lib.ref_target_arch.restype = C.c_char_p
    ref_lib.ref_max_threads.restype = _I
    ref_lib.ref_bench_gemm_batch.restype = C.c_double
    ref_lib.ref_bench_gemm_batch.argtypes = [_P, _P, _U, _I, _LL, _LL, _ULL, _P, _P, _P, _LL, _LL, _LL, _LL, _I, _P]
    ref_lib.ref_bench_fsspmdm.restype = C.c_double
    ref_lib.ref_bench_fsspmdm.argtypes = [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I]
    ref_lib.ref_bench_bcsc.restype = C.c_double
    ref_lib.ref_bench_bcsc.argtypes = [_P, _P, _U, _P, _P, _P, _P, _P, _I]


def iarr(*v):
    return (C.c_int * len(v))(*v)


def run_gemm(side, dims, types, flags, br_type, stride_a, stride_b, br, a, b, c, offs_a=None, offs_b=None, scf=0.0, mode=0):
    """side: `oracle` or `ref` dict. a/b: numpy arrays (or ctypes pointer arrays for address mode); c: numpy (in/out)."""
    def p(x):
        if x is None:
            return None
        if isinstance(x, np.ndarray):
            return x.ctypes.data
        return C.cast(x, C.c_void_p).value
    return side["gemm"](iarr(*dims), iarr(*types), flags, br_type, stride_a, stride_b, br, p(a), p(b), p(c), p(offs_a), p(offs_b), scf, mode)
