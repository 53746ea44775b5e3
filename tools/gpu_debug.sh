#!/bin/bash
mkdir -p gpurun_out
echo "=== tc tests"; timeout -s KILL 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "tcgen05 or multi_device or plan or pipeline" > gpurun_out/test_gemm_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/test_gemm_gpu.log
echo "=== sweep"; for pk in 1 0; do LIBXSMM_B200_TC_PACK=$pk timeout -s KILL 300 python bench.py --workload sweep --steps 10 > gpurun_out/bench_sweep_$pk.json 2> gpurun_out/bench_sweep_$pk.err; echo "pack $pk rc=$?"; tail -3 gpurun_out/bench_sweep_$pk.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_sweep_$pk.json'))
for p in d['points']:
    if p['type'].startswith('f16'): print({a:(round(b,4) if isinstance(b,float) else b) for a,b in p.items() if a in ('type','m','ms','hbm_frac','backend','error')})
PY
done
