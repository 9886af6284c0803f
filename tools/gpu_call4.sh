#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/debug_determinism.py > gpurun_out/determinism.log 2>&1
tail -n 12 gpurun_out/determinism.log
timeout 600 python tools/bench_linear.py > gpurun_out/bench_linear.log 2>&1
