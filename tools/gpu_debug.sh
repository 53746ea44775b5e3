#!/bin/bash
mkdir -p gpurun_out
echo "=== meltw tests"; timeout -s KILL 300 python -m pytest tests/test_meltw_gpu.py -m gpu -q -x > gpurun_out/test_meltw.log 2>&1; echo "rc=$?"; grep -n "AssertionError: (" gpurun_out/test_meltw.log | head -3 | cut -c1-400; tail -3 gpurun_out/test_meltw.log
echo "=== gemm tests"; timeout -s KILL 500 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/test_gemm.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/test_gemm.log
echo "=== mode R bench"; for ps in 1; do LIBXSMM_B200_TC_POOLSETS=$ps timeout -s KILL 300 python bench.py --workload brgemm_r --steps 10 > gpurun_out/bench_r_$ps.json 2> gpurun_out/bench_r.err; echo "poolsets $ps rc=$?"; tail -3 gpurun_out/bench_r.err; cut -c1-330 gpurun_out/bench_r_$ps.json | cut -c130-; done
echo "=== ts probe"; timeout -s KILL 300 python tools/ts_probe.py "" > gpurun_out/ts_probe.log 2>&1; cat gpurun_out/ts_probe.log
echo "=== bench"; timeout -s KILL 600 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
for k,v in d['also'].items():
    if 'points' in v:
        for p in v['points']: print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in p.items() if a in ('type','m','ms','hbm_frac','op','n','GBps','backend','error')})
    else: print(k, v.get('ms_per_step'), v.get('roofline',{}).get('frac'), v.get('error'))
PY
echo "=== ncu mode R"; timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:gemm_pool -c 1 -o gpurun_out/prof_tc_pool -f python bench.py --workload brgemm_r --steps 5 > gpurun_out/ncu_r.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_r.log
