#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu tests/test_fuzz_gpu.py tests/test_kernels_gpu.py -k "random" > gpurun_out/fuzz_bf16.log 2>&1; echo "bf16 rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/fuzz_bf16.log | head -40
DLB200_TEST_DTYPE=float16 timeout 900 python -m pytest -q -m gpu tests/test_fuzz_gpu.py tests/test_kernels_gpu.py -k "random" > gpurun_out/fuzz_fp16.log 2>&1; echo "fp16 rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/fuzz_fp16.log | head -40
