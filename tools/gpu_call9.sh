#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/full_size_parity.jsonl
( time timeout 900 python -m pytest tests -m gpu -q -x --durations=12 ) > gpurun_out/pytest_gpu_all.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_gpu_all.log
tail -n 30 gpurun_out/pytest_gpu_all.log | cut -c1-200
timeout 300 python tools/ncu_traffic.py > gpurun_out/ncu_traffic.log 2>&1; tail -n 3 gpurun_out/ncu_traffic.log
mkdir -p profiles; cp gpurun_out/ncu_traffic.json profiles/ncu_traffic.json 2>/dev/null
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file gpurun_out/racecheck_kernels.log python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "not full_size and not deadlock and not many_items and not bitwise_repeatable" > gpurun_out/racecheck_pytest.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/racecheck_pytest.log; tail -n 4 gpurun_out/racecheck_pytest.log; tail -n 3 gpurun_out/racecheck_kernels.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 7 --log-file gpurun_out/synccheck_kernels.log python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "not full_size and not deadlock and not many_items and not bitwise_repeatable" > gpurun_out/synccheck_pytest.log 2>&1
echo "synccheck rc=$?" >> gpurun_out/synccheck_pytest.log; tail -n 4 gpurun_out/synccheck_pytest.log; tail -n 3 gpurun_out/synccheck_kernels.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck_kernels.log python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/memcheck_pytest.log; tail -n 4 gpurun_out/memcheck_pytest.log; tail -n 3 gpurun_out/memcheck_kernels.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 200 python bench.py --config 3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 200 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 300 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
head -c 400 gpurun_out/bench_c2.json
