#!/bin/bash
# the GPU suites in float16, all failures listed (no -x)
mkdir -p gpurun_out
DLB200_TEST_DTYPE=float16 timeout 1500 python -m pytest -q -m gpu -k "not full_size" tests/test_kernels_gpu.py tests/test_model_gpu.py > gpurun_out/fp16_suite.log 2>&1; echo "rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/fp16_suite.log | head -60
