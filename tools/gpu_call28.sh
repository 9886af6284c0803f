#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:r[k] for k in ('value','ms_per_step')}, r['e2e']['value'], r['phases']['clip_prefill_ms'], r['phases']['decode_ms_per_token_step'])"; done
timeout 600 python tools/bench_linear.py 2>&1 | tail -25
