#!/bin/bash
# ncu --set full capture of the three BCSC kernel variants at the BASELINE size (run through gpurun; ~1 min each).
# Compare per variant: l1tex__m_xbar2l1tex_read_bytes (TMA volume), l1tex throughput, sm__pipe_tensor_cycles_active,
# smsp__average_warps_issue_stalled_* of the MMA warps, and the kernel duration.
mkdir -p gpurun_out
for v in 1 2 3; do
  export LIBXSMM_B200_BCSC_V2=0 LIBXSMM_B200_BCSC_V3=0
  [ $v = 2 ] && export LIBXSMM_B200_BCSC_V2=1
  [ $v = 3 ] && export LIBXSMM_B200_BCSC_V3=1
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:'bcsc_tc' -s 2 -c 1 -f -o gpurun_out/prof_bcsc_v$v \
    python tools/bcsc_probe.py bench "" > gpurun_out/ncu_bcsc_v$v.log 2>&1
  echo "variant $v rc=$?"
done
