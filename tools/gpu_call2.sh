#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "linear" > gpurun_out/linear_tests.log 2>&1
echo "rc=$?" >> gpurun_out/linear_tests.log
tail -15 gpurun_out/linear_tests.log
timeout 300 python tools/bench_linear.py > gpurun_out/bench_linear.log 2>&1
tail -30 gpurun_out/bench_linear.log | cut -c1-400
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_full_size_parity_gpu.py > gpurun_out/pytest_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
head -c 1500 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
