#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_full_size_parity_gpu.py > gpurun_out/pytest_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-200
for i in 1 2; do
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_$i.json 2> gpurun_out/bench_c2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c2_$i.json').read().strip().splitlines()[-1])
print('BENCH value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2), 'prefill', round(d['phases']['clip_prefill_ms'],2), 'decode_ms', round(d['phases']['decode_ms_per_token_step'],4))
PY
done
