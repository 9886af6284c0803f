#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_all.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_all.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
r=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print({k:r[k] for k in ('value','ms_per_step','gpu_launches')}, r['e2e']['value'], r['phases']['clip_prefill_ms'], r['phases']['decode_ms_per_token_step'], r['roofline']['frac'], r['roofline']['traffic'], r['cpu_baseline']['value'], r['clocks'])
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
