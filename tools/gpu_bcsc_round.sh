#!/bin/bash
# BCSC kernel round on the GPU box (through gpurun): parity of the tensor-core kernels, then timings per setting.
mkdir -p gpurun_out
echo "=== sparse tests"; timeout -s KILL 240 python -m pytest tests/test_sparse_gpu.py -m gpu -q -x > gpurun_out/test_sparse_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/test_sparse_gpu.log
echo "=== sparse tests (part-major)"; LIBXSMM_B200_BCSC_KMAJOR=0 timeout -s KILL 240 python -m pytest tests/test_sparse_gpu.py -m gpu -q -x -k bcsc > gpurun_out/test_sparse_gpu_pm.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/test_sparse_gpu_pm.log
echo "=== bench settings"
timeout -s KILL 240 python tools/bcsc_probe.py bench "" "PDL=1" "KMAJOR=0" "RESIDENT=0,KPC=4,BST=2" "MMAW=2" > gpurun_out/bcsc_bench.log 2>&1; echo "rc=$?"; cat gpurun_out/bcsc_bench.log | tail -12
echo "=== scale"; timeout -s KILL 300 python tools/bcsc_probe.py scale "" > gpurun_out/bcsc_scale.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bcsc_scale.log
echo "=== ncu full"; timeout -s KILL 240 ncu --set full --clock-control none --import-source on -k regex:bcsc_ts -s 2 -c 1 -f -o gpurun_out/prof_bcsc_ts python tools/bcsc_probe.py bench "" > gpurun_out/ncu_bcsc.log 2>&1; echo "rc=$?"
