"""Tuning sweep (not a test): time the BRGEMM tile kernel and the fsspmdm kernel under the env knobs."""
import os, sys, json, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
import libxsmm_b200 as X


class A: steps = 10; warmup = 3

X.libxsmm_b200_set_blocking(0)
pk = bench.peaks()
kernel, a, b, c, (sa, sb, sc) = bench.brgemm_setup(X, torch, bench.BATCH)
bytes_alg = float(bench.BATCH) * (sa + sb + sc)


def t_brgemm():
    def step():
        assert X.libxsmm_b200_gemm_batch_strided(kernel, a.data_ptr(), b.data_ptr(), c.data_ptr(), sa, sb, sc, bench.BR, bench.BATCH) == 0
    total, per = bench.time_steps(torch, step, 10, 3)
    X.check()
    ms = sorted(per)[len(per) // 2]
    return ms, bytes_alg / (ms * 1e-3) / 1e9

for ctas, stages, ev in itertools.product((2, 3, 4), (2, 3, 4, 6), (0,)):
    if ctas * stages > 13:
        continue
    os.environ["LIBXSMM_B200_TC_STAGES"] = str(stages); os.environ["LIBXSMM_B200_TC_CTAS"] = str(ctas); os.environ["LIBXSMM_B200_TC_EVICT_FIRST"] = str(ev)
    ms, gbs = t_brgemm()
    print("brgemm stages=%d ctas=%d evict=%d : %.3f ms %.0f GB/s (%.1f%%)" % (stages, ctas, ev, ms, gbs, 100 * gbs / pk["hbm_gbs"]), flush=True)
bench.brgemm_check(X, torch, a, b, c, (sa, sb, sc))
del a, b, c
torch.cuda.empty_cache()
for st, w in itertools.product((2, 3), (16, 32)):
    os.environ["LIBXSMM_B200_SREG_STAGES"] = str(st); os.environ["LIBXSMM_B200_SREG_WARPS"] = str(w)
    r = bench.also_fsspmdm(X, torch, pk, A)
    print("fsspmdm stages=%d warps=%d : %.3f ms %.0f GB/s (%.1f%%)" % (st, w, r["ms_per_step"], r["roofline"]["achieved"], 100 * r["roofline"]["frac"]), flush=True)
