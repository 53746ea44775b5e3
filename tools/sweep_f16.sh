#!/bin/bash
# f16 part of the sweep under a few TC kernel settings
for cfg in "2 4" "4 4" "4 3" "3 4"; do
  set -- $cfg
  echo "== TC_CTAS=$1 TC_STAGES=$2"
  LIBXSMM_B200_TC_CTAS=$1 LIBXSMM_B200_TC_STAGES=$2 bash tools/sweep_i8.sh 2>&1 | grep f16
done
