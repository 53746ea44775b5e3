"""Diagnostic (not a test): run single tiles through the tcgen05 kernel and print an error map vs the oracle.
Used on the GPU box to localise descriptor/layout mistakes in one round trip:  python tools/tc_probe.py [m n k br count]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import libxsmm_b200 as X
import cases, gen
from gpu_util import dev, dispatch, host
from oracle_ffi import oracle, run_gemm


def probe(m, n, k, br, count, ta=gen.BF16, tc=gen.F32, structured=False):
    case = cases.GemmCase(m, n, k, ta, ta, gen.F32, tc, flags=cases.FLAG_BETA_0, br_type=3 if br > 1 else 0, br=br,
                          lda=(m + 7) // 8 * 8, ldb=(k + 7) // 8 * 8)
    ops = cases.Operands(case, count=count)
    if structured:   # A = e_(i==kk) pattern: C = B^T-ish, shows index permutations directly
        a = np.zeros(case.size_a * case.br * count, dtype=np.float32).reshape(count * case.br, case.k, case.lda)
        for kk in range(min(m, k)):
            a[:, kk, kk] = 1.0
        ops.a = gen.f32_to_bf16_bits(a.ravel())
        b = np.zeros(case.size_b * case.br * count, dtype=np.float32).reshape(count * case.br, case.n, case.ldb)
        for j in range(n):
            for kk in range(k):
                b[:, j, kk] = (j * 1.0 + kk / 128.0) / case.br
        ops.b = gen.f32_to_bf16_bits(b.ravel())
    kernel = dispatch(case, ops)
    print("case", case, "backend", X.libxsmm_b200_kernel_backend(kernel))
    d_a, d_b, d_c = dev(ops.a), dev(ops.b), dev(ops.c0)
    rc = X.libxsmm_b200_gemm_batch_strided(kernel, d_a.data_ptr(), d_b.data_ptr(), d_c.data_ptr(), ops.tile_a, ops.tile_b, ops.tile_c, case.br, count)
    err = X.libxsmm_b200_sync()
    print("rc", rc, "sync", err, X.libxsmm_b200_last_error_string())
    got = gen.to_f64(host(d_c, gen.NP_OF[tc]), tc).reshape(count, case.n, case.ldc)[:, :, :m]
    want = gen.to_f64(cases.ref_result(oracle, case, ops, run_gemm), tc).reshape(count, case.n, case.ldc)[:, :, :m]
    print("normf_rel", gen.normf_rel(want, got), "max abs", np.abs(want - got).max())
    for t in range(min(count, 3)):
        e = np.abs(want[t] - got[t])
        print("tile", t, "err by 8x8 block (rows=n/8, cols=m/8):")
        for jb in range(0, n, 8):
            print(" ".join("%7.3f" % e[jb:jb + 8, ib:ib + 8].max() for ib in range(0, m, 8)))
        if structured:
            print("got[n=0..3, m=0..15]:"); print(np.round(got[t][:4, :16], 3))
            print("want[n=0..3, m=0..15]:"); print(np.round(want[t][:4, :16], 3))


if __name__ == "__main__":
    argv = [int(x) for x in sys.argv[1:6]]
    m, n, k, br, count = (argv + [64, 64, 64, 1, 1][len(argv):])
    probe(m, n, k, br, count)
    probe(m, n, k, br, count, structured=True)
