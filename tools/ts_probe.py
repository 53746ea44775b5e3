"""Diagnostic: milliseconds of the VNNI-A tensor-core GEMM (gemm_ts) per sweep point under environment settings.
usage: ts_probe.py "" "TS=0" "TS_CTAS=1,TS_STAGES=2" ...   (LIBXSMM_B200_ prefix implied; one subprocess per setting)
       ts_probe.py one <m>                                    (a single int8 point, for ncu)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_points(ms_list, kinds=("i8", "bf16")):
    import torch
    import libxsmm_b200 as X
    import bench
    I8, U8, I32, BF16, F32 = 12, 13, 8, 2, 1
    batch = 32768
    out = []
    for kind in kinds:
        for m in ms_list:
            if kind == "i8":
                ta, tb, tcc, tcomp, esz, csz = U8, I8, I32, I32, 1, 4
            else:
                ta, tb, tcc, tcomp, esz, csz = BF16, BF16, F32, F32, 2, 4
            shape = X.libxsmm_create_gemm_shape(m, m, m, m, m, m, ta, tb, tcc, tcomp)
            kernel = X.libxsmm_dispatch_gemm(shape, bench.FLAG_BETA_0 | X.GEMM_FLAG_VNNI_A, 0)
            if not kernel:
                out.append("%s m=%d NULL" % (kind, m)); continue
            a = torch.randint(0, 5, (batch * m * m * esz,), dtype=torch.uint8, device="cuda")
            b = torch.randint(0, 5, (batch * m * m * esz,), dtype=torch.uint8, device="cuda")
            c = torch.empty(batch * m * m * csz, dtype=torch.uint8, device="cuda")
            sa = sb = m * m * esz; sc = m * m * csz

            def step():
                assert X.libxsmm_b200_gemm_batch_strided(kernel, a.data_ptr(), b.data_ptr(), c.data_ptr(), sa, sb, sc, 1, batch) == 0
            step(); X.check()
            _, per = bench.time_steps(torch, step, 6, 2)
            ms = sorted(per)[len(per) // 2]
            out.append("%s m=%d be=%d %.4f ms (%.0f GB/s)" % (kind, m, int(X.libxsmm_b200_kernel_backend(kernel)), ms, batch * (sa + sb + sc) / ms / 1e6))
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        print(run_points([int(sys.argv[2])], kinds=(sys.argv[3] if len(sys.argv) > 3 else "i8",)))
    elif len(sys.argv) > 1 and sys.argv[1] == "child":
        print(" | ".join(run_points([int(x) for x in os.environ.get("TS_PROBE_M", "16,32,64,128").split(",")])), flush=True)
    else:
        for cfg in (sys.argv[1:] or [""]):
            env = dict(os.environ)
            for kv in filter(None, cfg.split(",")):
                k, v = kv.split("="); env["LIBXSMM_B200_" + k] = v
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True, timeout=200)
            print("settings %-26s %s" % (cfg or "(default)", r.stdout.strip() or r.stderr[-400:]), flush=True)
