#!/bin/bash
mkdir -p gpurun_out
echo "=== gemm tests"; timeout -s KILL 500 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/test_gemm.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/test_gemm.log
echo "=== meltw tests"; timeout -s KILL 300 python -m pytest tests/test_meltw_gpu.py -m gpu -q -x > gpurun_out/test_meltw.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/test_meltw.log
echo "=== ts probe"; timeout -s KILL 300 python tools/ts_probe.py "" "TS_PACK=0" "TS_GRP=4" "TS_GRP=2" > gpurun_out/ts_probe.log 2>&1; cat gpurun_out/ts_probe.log
echo "=== mode R bench"; timeout -s KILL 300 python bench.py --workload brgemm_r --steps 10 > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; echo "rc=$?"; tail -3 gpurun_out/bench_r.err; cut -c1-400 gpurun_out/bench_r.json
echo "=== sweep"; timeout -s KILL 300 python bench.py --workload sweep --steps 10 > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err; echo "rc=$?"; tail -3 gpurun_out/bench_sweep.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_sweep.json'))
for p in d['points']: print({a:(round(b,4) if isinstance(b,float) else b) for a,b in p.items() if a in ('type','m','ms','hbm_frac','backend','error')})
PY
echo "=== ncu mode R"; timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:gemm_tc -c 1 -o gpurun_out/prof_tc_pool -f python bench.py --workload brgemm_r --steps 5 > gpurun_out/ncu_r.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_r.log
