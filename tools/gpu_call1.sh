#!/bin/bash
# round 2, call 1: existing GPU tests, full-size oracle parity, bench configs 2-5, memcheck of the kernel tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/full_size_parity.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_full_size_parity_gpu.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 1500 python -m pytest tests/test_full_size_parity_gpu.py -m gpu -q -s > gpurun_out/parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/parity.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 300 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 300 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 400 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck_kernels.log python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/memcheck_pytest.log
tail -5 gpurun_out/pytest_gpu.log gpurun_out/parity.log gpurun_out/memcheck_pytest.log
head -c 600 gpurun_out/bench_c2.json
