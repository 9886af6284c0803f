#!/bin/bash
timeout 600 python -m pytest -q -m gpu tests/test_kernels_gpu.py -k "back_to_back or linear" 2>&1 | tail -3
timeout 900 python -m pytest -q -m gpu tests/test_pdl_bitwise_gpu.py tests/test_full_size_parity_gpu.py -k "pdl or batch_sizes" 2>&1 | tail -3
timeout 600 python tools/bench_linear.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print(f\"{r['shape'][:46]:46s} SK {r['ours_us']:7.1f}  whole {r['ours_whole_tiles_us']:7.1f}  lib {r['library_path_us']:7.1f}\")"
