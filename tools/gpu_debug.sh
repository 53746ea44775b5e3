#!/bin/bash
mkdir -p gpurun_out
echo "=== gemm tests"; timeout -s KILL 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/test_gemm_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/test_gemm_gpu.log
echo "=== golden"; timeout -s KILL 300 python -m pytest tests/test_golden.py -m gpu -q -x > gpurun_out/test_golden.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/test_golden.log
echo "=== ncu launches headline"; timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_headline.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-also > gpurun_out/ncu_launches_h.log 2>&1; echo "rc=$?"
echo "=== ncu launches mode R"; timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r.csv python bench.py --workload brgemm_r --steps 5 > gpurun_out/ncu_launches_r.log 2>&1; echo "rc=$?"
echo "=== ncu full mode R"; timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gemm_pool -s 2 -c 1 -f -o gpurun_out/prof_tc_pool python bench.py --workload brgemm_r --steps 3 > gpurun_out/ncu_r.log 2>&1; echo "rc=$?"
echo "=== mode R"; timeout -s KILL 300 python bench.py --workload brgemm_r --steps 10 > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; echo "rc=$?"; cut -c130-330 gpurun_out/bench_r.json
