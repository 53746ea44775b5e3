#!/bin/bash
# int8 part of the sweep only (quick)
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, bench
import libxsmm_b200 as X
X.libxsmm_b200_set_stream(torch.cuda.current_stream().cuda_stream); X.libxsmm_b200_set_blocking(0)
class A: steps = 10; warmup = 3
r = bench.sweep(X, torch, bench.peaks(), A)
for p in r["points"]:
    if "gflops" in p: print(p["type"], p["m"], "%.0f GF/s %.4f ms %.0f GB/s frac %.3f backend %d" % (p["gflops"], p["ms"], p["gbs"], p["hbm_frac"], p["backend"]))
PY
