#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "dec_gemm or embed_rows_ssq or inplace" > gpurun_out/decgemm_tests.log 2>&1
echo "rc=$?" >> gpurun_out/decgemm_tests.log
tail -n 6 gpurun_out/decgemm_tests.log | cut -c1-250
timeout 60 python tools/bench_decode_ops.py 2>&1 | tail -n 14
for chain in v2 v1 v2; do
  DLB200_DECODE_CHAIN=$chain timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$chain.json 2> gpurun_out/bench_$chain.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$chain.json').read().strip().splitlines()[-1])
    print('CHAIN', '$chain', 'value', round(d['value'],1), 'decode_ms', round(d['phases']['decode_ms_per_token_step'],4), 'prefill', round(d['phases']['clip_prefill_ms'],2), 'launches', d['gpu_launches'])
except Exception as e:
    print('CHAIN $chain failed', e); print(open('gpurun_out/bench_$chain.err').read()[-800:])
PY
done
timeout 200 python -m pytest tests/test_model_gpu.py -m gpu -x -q > gpurun_out/pytest_model.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_model.log
tail -n 8 gpurun_out/pytest_model.log | cut -c1-250
