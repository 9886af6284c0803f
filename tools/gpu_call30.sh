#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ncu_traffic.py > gpurun_out/ncu_traffic.log 2>&1; echo "ncu_traffic rc=$?"
cp gpurun_out/ncu_traffic.json profiles/ncu_traffic.json
for c in 2 3 4 5; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_config${c}_v3.json 2> gpurun_out/bench_config${c}_v3.err; echo "bench config $c rc=$?"
  python - <<PY
import json
r=json.loads(open('gpurun_out/bench_config${c}_v3.json').read().strip().splitlines()[-1])
print($c, {k:r[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', round(r['e2e']['value'],1), 'prefill', round(r['phases']['clip_prefill_ms'],2), 'decode', round(r['phases']['decode_ms_per_token_step'],3), 'frac', round(r['roofline']['frac'],3), 'traffic', r['roofline']['traffic'], 'cpu', r.get('cpu_baseline',{}).get('value'))
PY
done
