#!/bin/bash
mkdir -p gpurun_out
echo "=== 2-GPU bench"; timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?"; tail -4 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_2gpu.json'))
    print('n', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'])
    print('strong', json.dumps(d.get('strong'))[:1500])
    for k,v in d.get('also',{}).items(): print(k, v.get('ms_per_step'), v.get('roofline',{}).get('frac'), v.get('error'))
except Exception as e: print('parse failed', e); print(open('gpurun_out/bench_2gpu.json').read()[:600])
PY
echo "=== 2-GPU reference arm"; timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_ref_2gpu.json; tail -3 gpurun_out/bench_ref_2gpu.err
echo "=== meltw tests"; timeout -s KILL 300 python -m pytest tests/test_meltw_gpu.py -m gpu -q -x > gpurun_out/test_meltw.log 2>&1; echo "rc=$?"; grep -n "AssertionError" gpurun_out/test_meltw.log | head -3 | cut -c1-300; tail -2 gpurun_out/test_meltw.log
echo "=== pool tests"; timeout -s KILL 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "pool" > gpurun_out/test_gemm.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/test_gemm.log
echo "=== mode R"; timeout -s KILL 300 python bench.py --workload brgemm_r --steps 10 > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; echo "rc=$?"; cut -c130-330 gpurun_out/bench_r.json
