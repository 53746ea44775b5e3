#!/bin/bash
mkdir -p gpurun_out
echo "=== ts gemm"; timeout -s KILL 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "vnni_a_on" > gpurun_out/test_ts.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/test_ts.log
for pdl in 1 0; do
echo "=== bcsc tests PDL=$pdl"; LIBXSMM_B200_BCSC_PDL=$pdl timeout -s KILL 200 python -m pytest tests/test_sparse_gpu.py -m gpu -q -k "bcsc" > gpurun_out/test_bcsc_pdl$pdl.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/test_bcsc_pdl$pdl.log
done
echo "=== probe PDL=1"; timeout -s KILL 200 python tools/bcsc_probe.py > gpurun_out/probe1.log 2>&1; tail -40 gpurun_out/probe1.log
