"""Diagnostic: BCSC tensor-core kernel vs oracle with an error map per (block-column, m_block)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import libxsmm_b200 as X
import cases, gen
from gpu_util import dev, host
from oracle_ffi import oracle
from test_oracle_vs_ref import _bcsc_inputs, _run_bcsc


def probe(mblocks, M, K, N, bk, bn, dens, beta0=1, seed=0):
    rng = np.random.default_rng(seed)
    ta = gen.BF16
    a, bvals, colptr, rowidx, c0 = _bcsc_inputs(rng, ta, ta, ta, mblocks, M, K, N, bk, bn, dens)
    flags = (cases.FLAG_BETA_0 if beta0 else 0) | cases.FLAG_VNNI_A
    sh = X.libxsmm_create_gemm_shape(mblocks, 0, K, K, 0, N, ta, ta, ta, gen.F32)
    k = X.libxsmm_create_packed_spgemm_bcsc(sh, flags, 0, X.SpgemmConfig(M, bk, bn))
    d_a, d_b, d_cp, d_ri, d_c = dev(a), dev(bvals), dev(colptr), dev(rowidx), dev(c0)
    X.call_gemm(k, d_a, d_b, d_c, colptr=d_cp, rowidx=d_ri, nblocks=N // bn)
    rc = X.libxsmm_b200_sync()
    got = gen.to_f64(host(d_c, np.uint16), ta).reshape(mblocks, N, M)
    want = c0.copy(); _run_bcsc(oracle, (ta, ta, gen.F32, ta), (mblocks, M, K, N, bk, bn), flags, a, bvals, colptr, rowidx, want)
    want = gen.to_f64(want, ta).reshape(mblocks, N, M)
    print("cfg", (mblocks, M, K, N, bk, bn, dens, beta0), "variant", X.libxsmm_b200_bcsc_variant(k, N // bn), "sync", rc, X.libxsmm_b200_last_error_string(), "normf_rel %.3e" % gen.normf_rel(want, got), flush=True)
    e = np.abs(want - got)
    for mb in range(min(mblocks, 6)):
        print(" mb%d per block-col max err:" % mb, " ".join("%.2f" % e[mb, j * bn:(j + 1) * bn].max() for j in range(N // bn)))


if __name__ == "__main__" and len(sys.argv) == 1:
    probe(4, 32, 64, 64, 32, 32, 1.0)
    probe(4, 32, 128, 64, 32, 32, 0.5)
    probe(5, 32, 512, 512, 32, 32, 0.5, beta0=0)
    probe(3, 16, 64, 96, 16, 32, 0.5)
    probe(2, 64, 256, 128, 32, 16, 0.5)
    probe(3, 32, 128, 128, 64, 32, 0.5)
    probe(9, 32, 96, 320, 32, 32, 0.4)
    probe(3, 128, 160, 64, 16, 16, 0.6, beta0=0)
    probe(5, 32, 512, 512, 32, 32, 0.5, beta0=0, seed=3)


if len(sys.argv) > 1 and sys.argv[1] == "bench":
    # plain milliseconds per setting, e.g. "KPC=1,BST=6" (LIBXSMM_B200_BCSC_ prefix implied)
    import torch
    import bench

    class A: steps = 8; warmup = 3; no_cpu = True
    for cfg in (sys.argv[2:] or [""]):
        for kv in filter(None, cfg.split(",")):
            k, v = kv.split("="); os.environ["LIBXSMM_B200_BCSC_" + k] = v
        r = bench.also_bcsc(X, torch, bench.peaks(), A)
        print("settings %-28s bcsc ms %.4f  frac %.3f" % (cfg or "(default)", r["ms_per_step"], r["roofline"]["frac"]), flush=True)
        for kv in filter(None, cfg.split(",")):
            os.environ.pop("LIBXSMM_B200_BCSC_" + kv.split("=")[0], None)


if len(sys.argv) > 1 and sys.argv[1] == "scale":
    # milliseconds against the number of m_blocks (148 SMs x 4 m_blocks per group): separates fixed from per-item cost
    import torch
    import bench

    class A: steps = 12; warmup = 3; no_cpu = True
    for cfg in (sys.argv[2:] or [""]):
        for kv in filter(None, cfg.split(",")):
            k, v = kv.split("="); os.environ["LIBXSMM_B200_BCSC_" + k] = v
        out = []
        for groups_per_cta in (1, 2, 4, 8, 14):
            r = bench.also_bcsc(X, torch, bench.peaks(), A, mblocks=148 * 4 * groups_per_cta)
            out.append("%d:%.4f" % (groups_per_cta, r["ms_per_step"]))
        print("settings %-24s ms by groups/CTA  %s" % (cfg or "(default)", "  ".join(out)), flush=True)
        for kv in filter(None, cfg.split(",")):
            os.environ.pop("LIBXSMM_B200_BCSC_" + kv.split("=")[0], None)
