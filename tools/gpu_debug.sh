#!/bin/bash
mkdir -p gpurun_out
echo "=== gemm tests (pool, tc)"; timeout -s KILL 500 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "pool or tcgen05 or tensor" > gpurun_out/test_gemm.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/test_gemm.log
echo "=== meltw tests"; timeout -s KILL 300 python -m pytest tests/test_meltw_gpu.py -m gpu -q -x > gpurun_out/test_meltw.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/test_meltw.log
echo "=== mode R bench"; for pr in 1 0; do LIBXSMM_B200_TC_PAIR=$pr timeout -s KILL 300 python bench.py --workload brgemm_r --steps 10 > gpurun_out/bench_r_$pr.json 2> gpurun_out/bench_r.err; echo "pair $pr rc=$?"; tail -3 gpurun_out/bench_r.err; cut -c1-330 gpurun_out/bench_r_$pr.json | cut -c130-; done
echo "=== ncu mode R"; timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:gemm_tc -c 1 -o gpurun_out/prof_tc_pool -f python bench.py --workload brgemm_r --steps 5 > gpurun_out/ncu_r.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_r.log
