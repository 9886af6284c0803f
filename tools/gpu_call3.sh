#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 4 -c 2 -o gpurun_out/ncu_gemm2_oproj python tools/bench_linear.py --only "o_proj+res, layers 2-31" --no-lib --iters 2 --out gpurun_out/tmp.json > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 4 -c 1 -o gpurun_out/ncu_gemm2_qkv python tools/bench_linear.py --only "qkv, layers 0-1 (M=8*607)" --no-lib --iters 2 --out gpurun_out/tmp.json > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
timeout 600 python tools/bench_linear.py --quick > gpurun_out/bench_linear.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv >> gpurun_out/smi.txt 2>&1
cat gpurun_out/smi.txt
