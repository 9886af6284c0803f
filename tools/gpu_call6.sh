#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for mb in -1 -12 -24 -48 -1 -24; do
  DLB200_L2_PREFETCH_MB=$mb timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_l2pf_$mb.json 2> gpurun_out/bench_l2pf.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_l2pf_$mb.json').read().strip().splitlines()[-1])
print('L2PF', $mb, 'value', round(d['value'],1), 'decode_ms', round(d['phases']['decode_ms_per_token_step'],4), 'gemm_ms', round(d['phases']['weight_gemm_ms_per_decode_step'],4), 'prefill', round(d['phases']['clip_prefill_ms'],2))
PY
done
