#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for shape in "qkv, layers 0-1 (M=8*607)" "qkv, layers 2-31" "o_proj+res, layers 2-31" "gate|up+swiglu, layers 0-1" "CLIP fc1"; do
  tag=$(echo "$shape" | tr -c 'a-zA-Z0-9' '_' | cut -c1-24)
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 4 -c 1 -f -o gpurun_out/ncu_gemm2_final_$tag python tools/bench_linear.py --only "$shape" --no-lib --iters 2 --out gpurun_out/tmp.json > gpurun_out/ncu_g.log 2>&1
  tail -n 1 gpurun_out/ncu_g.log | cut -c1-120
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_prefill_final.csv python bench.py --steps 1 --warmup 1 --new-tokens 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_prefill.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 800 --csv --log-file gpurun_out/launches_decode_final.csv python bench.py --steps 1 --warmup 1 --new-tokens 6 --no-cpu-baseline --no-graph > gpurun_out/ncu_decode.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -n 8
