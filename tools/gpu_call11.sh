#!/bin/bash
# experiment: shared-memory footprints that let the o_proj GEMM's CTAs become resident (and prefetch Wo) while the decode attention runs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() {
  python -m dynamic_llava_b200.build > /dev/null 2>&1
  timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_exp.json 2> gpurun_out/bench_exp.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_exp.json').read().strip().splitlines()[-1])
    print('EXP $1', 'value', round(d['value'],1), 'decode_ms', round(d['phases']['decode_ms_per_token_step'],4), 'gemm_ms', round(d['phases']['weight_gemm_ms_per_decode_step'],4))
except Exception as e:
    print('EXP $1 failed', e); print(open('gpurun_out/bench_exp.err').read()[-600:])
PY
}
run "baseline(AD3,SG5)"
sed -i 's/#define AD_STAGES 3/#define AD_STAGES 2/' dynamic_llava_b200/csrc/attn_decode.cu
run "AD2,SG5"
sed -i 's/#define SG_STAGES 5 /#define SG_STAGES 4 /' dynamic_llava_b200/csrc/skinny_gemm.cu
run "AD2,SG4"
sed -i 's/#define AD_STAGES 2/#define AD_STAGES 3/' dynamic_llava_b200/csrc/attn_decode.cu
run "AD3,SG4"
