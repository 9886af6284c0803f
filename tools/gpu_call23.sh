#!/bin/bash
for m in 0 2 4 8 6; do echo "--- DLB200_NO_PDL=$m"; DLB200_NO_PDL=$m timeout 300 python tools/debug_b_sweep.py 2 3 2>&1 | grep -E "^(OK|FAIL)" | cut -c1-60; done
