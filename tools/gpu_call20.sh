#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu tests/test_kernels_gpu.py -k "select_compact" 2>&1 | tail -5
timeout 600 python -m pytest -q -m gpu tests/test_fuzz_gpu.py tests/test_model_gpu.py -k "not full_size" 2>&1 | tail -3
timeout 600 python tools/k1_sweep.py 2>&1 | tail -8
