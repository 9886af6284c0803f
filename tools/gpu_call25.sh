#!/bin/bash
timeout 1400 python tools/full_size_sweep.py 20 16 2>&1 | grep -E "^(OK|FAIL|[0-9]+/)"
DLB200_TEST_DTYPE=float16 timeout 900 python tools/full_size_sweep.py 8 100 2>&1 | grep -E "^(OK|FAIL|[0-9]+/)"
