#!/bin/bash
mkdir -p gpurun_out
echo "=== pool tests"; timeout -s KILL 400 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "pool" > gpurun_out/test_gemm_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/test_gemm_gpu.log
echo "=== mode R"; for is in 2 1 2 1; do LIBXSMM_B200_TC_POOL_ISSUERS=$is timeout -s KILL 300 python bench.py --workload brgemm_r --steps 10 > gpurun_out/bench_r_$is.json 2> gpurun_out/bench_r.err; echo "issuers $is rc=$?"; tail -2 gpurun_out/bench_r.err; cut -c130-250 gpurun_out/bench_r_$is.json; done
echo "=== ncu mode R"; timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gemm_pool -s 2 -c 1 -f -o gpurun_out/prof_tc_pool2 python bench.py --workload brgemm_r --steps 3 > gpurun_out/ncu_r.log 2>&1; echo "rc=$?"
