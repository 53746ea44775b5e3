#!/bin/bash
mkdir -p gpurun_out
echo "=== tc tests"; timeout -s KILL 400 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "tcgen05" > gpurun_out/test_gemm_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/test_gemm_gpu.log
echo "=== headline"; timeout -s KILL 300 python bench.py --no-also --no-cpu --no-e2e --steps 20 > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "rc=$?"; tail -2 gpurun_out/bench_h.err; python -c "
import json; d=json.load(open('gpurun_out/bench_h.json')); print('headline', d['ms_per_step'], d['roofline']['frac'])"
echo "=== sweep"; timeout -s KILL 300 python bench.py --workload sweep --steps 10 > gpurun_out/bench_sweep.json 2> gpurun_out/bench_sweep.err; echo "rc=$?"; tail -2 gpurun_out/bench_sweep.err; python -c "
import json; d=json.load(open('gpurun_out/bench_sweep.json'))
for p in d['points']: print(p['type'], p['m'], round(p['ms'],4), round(p['hbm_frac'],4))"
