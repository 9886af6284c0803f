#!/bin/bash
timeout 600 python -m pytest -q -m gpu tests/test_kernels_gpu.py -k "back_to_back or linear" 2>&1 | tail -4
timeout 600 python tools/debug_b_sweep.py 1 2 3 4 6 7 2>&1 | grep -E "^(OK|FAIL)" | cut -c1-80
