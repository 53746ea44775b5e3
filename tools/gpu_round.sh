#!/bin/bash
# executed on the GPU box through gpurun; everything interesting lands in gpurun_out/. Every step has its own short timeout.
mkdir -p gpurun_out
echo "=== smoke"; timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
for f in test_sparse_gpu test_meltw_gpu test_golden test_ref_drivers test_meqn test_gemm_gpu test_c_relink; do
  echo "=== $f"; timeout -s KILL 300 python -m pytest tests/$f.py -m gpu -q -x > gpurun_out/$f.log 2>&1; echo "$f rc=$?"; tail -8 gpurun_out/$f.log
done
if [ "$1" != "nobench" ]; then
echo "=== bench"; timeout -s KILL 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "=== bench reference arm"; timeout -s KILL 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
fi
if [ "$1" = "ncu" ]; then
echo "=== ncu launches"; timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_t|sreg|bcsc|gemm_simt|gemm_i8|meltw|packed' -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
echo "=== ncu full tc"; timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/prof_tc python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-also > gpurun_out/ncu_tc.log 2>&1; echo "rc=$?"
echo "=== ncu full sreg"; timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:sreg_kernel -s 2 -c 1 -f -o gpurun_out/prof_sreg python bench.py --workload fsspmdm --steps 3 --warmup 3 > gpurun_out/ncu_sreg.log 2>&1; echo "rc=$?"
echo "=== ncu full bcsc"; timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:bcsc_t -s 2 -c 1 -f -o gpurun_out/prof_bcsc python bench.py --workload bcsc --steps 3 --warmup 3 > gpurun_out/ncu_bcsc.log 2>&1; echo "rc=$?"
echo "=== ncu full mode R"; timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o gpurun_out/prof_tc_pool python bench.py --workload brgemm_r --steps 3 > gpurun_out/ncu_r.log 2>&1; echo "rc=$?"
echo "=== ncu full gemm_ts (int8 64^3)"; timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:gemm_ts -s 1 -c 1 -f -o gpurun_out/prof_ts_i8_64 python tools/ts_probe.py one 64 > gpurun_out/ncu_ts.log 2>&1; echo "rc=$?"
fi
ls -la gpurun_out | head -40
