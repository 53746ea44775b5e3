#!/bin/bash
# BCSC kernel round on the GPU box (through gpurun): parity of both tensor-core kernels, then timings per setting.
mkdir -p gpurun_out
echo "=== sparse tests"; timeout -s KILL 900 python -m pytest tests/test_sparse_gpu.py -m gpu -q -x -k bcsc > gpurun_out/test_sparse_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/test_sparse_gpu.log
echo "=== sparse tests (streamed B)"; LIBXSMM_B200_BCSC_RESIDENT=0 timeout -s KILL 900 python -m pytest tests/test_sparse_gpu.py -m gpu -q -x -k bcsc > gpurun_out/test_sparse_gpu_stream.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/test_sparse_gpu_stream.log
echo "=== bench settings"
timeout -s KILL 600 python tools/bcsc_probe.py bench "" "V1=1" "RESIDENT=0" "RESIDENT=0,KPC=4,BST=2" "RAW=3" "RAW=4" "RAW=5" "MMAW=2" "MMAW=1" > gpurun_out/bcsc_bench.log 2>&1; echo "rc=$?"; cat gpurun_out/bcsc_bench.log | tail -12
echo "=== scale"; timeout -s KILL 300 python tools/bcsc_probe.py scale "" > gpurun_out/bcsc_scale.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/bcsc_scale.log
echo "=== launch list"; timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'bcsc' -c 30 --csv --log-file gpurun_out/bcsc_launches.csv python tools/bcsc_probe.py bench "" > gpurun_out/bcsc_ncu_launches.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/bcsc_launches.csv | cut -c1-200
