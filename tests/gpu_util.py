"""helpers for the -m gpu tests: device buffers are torch byte tensors, everything goes through the C ABI."""
import ctypes as C

import numpy as np
import torch

import libxsmm_b200 as X


def dev(arr):
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).copy()).cuda()


def host(t, npdtype):
    return t.cpu().numpy().view(npdtype).copy()


def shape_of(case):
    return X.libxsmm_create_gemm_shape(case.m, case.n, case.k, case.lda, case.ldb, case.ldc, case.ta, case.tb, case.tc, case.tcomp)


def dispatch(case, ops):
    """libxsmm_dispatch_gemm / _brgemm for a cases.GemmCase; returns the handle (int address) or None."""
    sh = shape_of(case)
    if case.br_type == 0:
        return X.libxsmm_dispatch_gemm(sh, case.flags, 0)
    brt = {1: X.GEMM_BATCH_REDUCE_ADDRESS, 2: X.GEMM_BATCH_REDUCE_OFFSET, 3: X.GEMM_BATCH_REDUCE_STRIDE}[case.br_type]
    cfg = X.libxsmm_create_gemm_batch_reduce_config(brt, ops.stride_a if ops is not None else 0, ops.stride_b if ops is not None else 0, 0)
    return X.libxsmm_dispatch_brgemm(sh, case.flags, 0, cfg)


def run_single_calls(kernel, case, ops, d_a, d_b, d_c):
    """one handle call per tile, exactly like the reference drivers' loops"""
    keep = []
    fn = X.GEMMFUNCTION(kernel)
    for t in range(ops.count):
        p = X.GemmParam()
        br = C.c_ulonglong(case.br); keep.append(br)
        p.op.tertiary = C.addressof(br)
        p.c.primary = d_c.data_ptr() + t * ops.tile_c
        if case.br_type == 1:
            ops.case_br = case.br
            aa, ab = ops.addr_arrays(d_a.data_ptr(), d_b.data_ptr(), t); keep += [aa, ab]
            p.a.primary, p.b.primary = C.addressof(aa), C.addressof(ab)
        else:
            p.a.primary, p.b.primary = d_a.data_ptr() + t * ops.tile_a, d_b.data_ptr() + t * ops.tile_b
        if ops.offs_a is not None:
            p.a.secondary, p.b.secondary = ops.offs_a.ctypes.data, ops.offs_b.ctypes.data
        if ops.scf:
            s = C.c_float(ops.scf); keep.append(s); p.c.tertiary = C.addressof(s)
        fn(C.byref(p))
    X.check()
    return keep
