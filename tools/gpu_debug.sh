#!/bin/bash
mkdir -p gpurun_out
echo "=== pooled + ts tests"; timeout -s KILL 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "address_mode_pool or vnni_a_on" > gpurun_out/test_pool.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/test_pool.log
echo "=== ts probe"; timeout -s KILL 400 python tools/ts_probe.py "" "TS=0" "TS_CTAS=1" "TS_CTAS=2" "TS_STAGES=2" > gpurun_out/ts_probe.log 2>&1; cat gpurun_out/ts_probe.log
echo "=== mode R bench"; timeout -s KILL 300 python bench.py --workload brgemm_r --steps 10 > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; echo "rc=$?"; tail -3 gpurun_out/bench_r.err; cut -c1-900 gpurun_out/bench_r.json
echo "=== ncu ts"; timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:gemm_ts -c 1 -o gpurun_out/prof_ts_i8_64 -f python tools/ts_probe.py one 64 > gpurun_out/ncu_ts.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_ts.log
