#!/bin/bash
# ncu captures for profiles/ (run through gpurun, one GPU). Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
echo "=== launch list"; timeout -s KILL 420 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc|sreg|bcsc|gemm_i8|gemm_simt' -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
echo "=== full bcsc_tc"; timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:bcsc_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_bcsc python tools/bcsc_probe.py bench "" > gpurun_out/ncu_bcsc.log 2>&1; echo "rc=$?"
ls -la gpurun_out | head -20
