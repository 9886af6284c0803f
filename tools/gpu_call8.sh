#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:dec_gemm_kernel<\(int\)16, \(int\)1, \(int\)2>' -s 30 -c 1 -o gpurun_out/ncu_dec_gemm_gu python tools/bench_decode_ops.py > gpurun_out/ncu_dg.log 2>&1
tail -n 3 gpurun_out/ncu_dg.log
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:skinny_gemm_kernel<\(int\)16, \(int\)1>' -s 30 -c 1 -o gpurun_out/ncu_skinny_gu python tools/bench_decode_ops.py > gpurun_out/ncu_sg.log 2>&1
tail -n 3 gpurun_out/ncu_sg.log
