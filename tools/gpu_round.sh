#!/bin/bash
# executed on the GPU box through gpurun; everything interesting lands in gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | grep -E "Model name|^CPU\(s\)|Flags" | cut -c1-400 >> gpurun_out/nproc.txt
echo "=== probe"; timeout -s KILL 300 python tools/tc_probe.py 64 64 64 1 1 > gpurun_out/probe.log 2>&1; echo "probe rc=$?"; head -40 gpurun_out/probe.log
echo "=== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
for f in test_gemm_gpu test_sparse_gpu test_meltw_gpu; do
  echo "=== $f"; timeout -s KILL 900 python -m pytest tests/$f.py -m gpu -q -x --timeout 600 > gpurun_out/$f.log 2>&1; echo "$f rc=$?"; tail -25 gpurun_out/$f.log
done
echo "=== bench"; timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "=== ncu launches"; timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
echo "=== ncu full tc"; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/prof_tc python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-also > gpurun_out/ncu_tc.log 2>&1; echo "rc=$?"
echo "=== ncu full sreg"; timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:sreg_kernel -s 2 -c 1 -f -o gpurun_out/prof_sreg python bench.py --workload fsspmdm --steps 3 --warmup 3 > gpurun_out/ncu_sreg.log 2>&1; echo "rc=$?"
ls -la gpurun_out
