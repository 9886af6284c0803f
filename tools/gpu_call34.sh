#!/bin/bash
# sanitizer pass over the final build: memcheck on ALL kernel tests (bf16) and on the fp16 library's kernel tests; racecheck + synccheck on the rewritten K1
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck_kernels_v3.log python -m pytest tests/test_kernels_gpu.py -m gpu -q > gpurun_out/memcheck_pytest_v3.log 2>&1
echo "memcheck bf16 rc=$?"; tail -n 2 gpurun_out/memcheck_pytest_v3.log; tail -n 2 gpurun_out/memcheck_kernels_v3.log
DLB200_TEST_DTYPE=float16 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck_kernels_fp16.log python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "not random" > gpurun_out/memcheck_pytest_fp16.log 2>&1
echo "memcheck fp16 rc=$?"; tail -n 2 gpurun_out/memcheck_pytest_fp16.log; tail -n 2 gpurun_out/memcheck_kernels_fp16.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 --log-file gpurun_out/racecheck_k1.log python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "select_compact" > gpurun_out/racecheck_k1_pytest.log 2>&1
echo "racecheck K1 rc=$?"; tail -n 2 gpurun_out/racecheck_k1_pytest.log; grep -c "Race reported\|hazard" gpurun_out/racecheck_k1.log; tail -n 3 gpurun_out/racecheck_k1.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 --log-file gpurun_out/synccheck_k1.log python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "select_compact" > gpurun_out/synccheck_k1_pytest.log 2>&1
echo "synccheck K1 rc=$?"; tail -n 2 gpurun_out/synccheck_k1_pytest.log; tail -n 3 gpurun_out/synccheck_k1.log
