#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "linear" > gpurun_out/linear_tests.log 2>&1
echo "rc=$?" >> gpurun_out/linear_tests.log
tail -n 15 gpurun_out/linear_tests.log
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_full_size_parity_gpu.py > gpurun_out/pytest_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
tail -n 5 gpurun_out/bench_c2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_prefill.csv python bench.py --steps 1 --warmup 1 --new-tokens 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_prefill.log 2>&1
tail -n 3 gpurun_out/ncu_prefill.log | cut -c1-300
