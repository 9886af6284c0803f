#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/full_size_parity.jsonl
timeout 1500 python -m pytest tests/test_full_size_parity_gpu.py -m gpu -q -x -k "toy or config5 or config4 or ragged" > gpurun_out/parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/parity.log
tail -n 5 gpurun_out/parity.log
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -k "nocache_properties or full_size" > gpurun_out/pytest_gpu2.log 2>&1
tail -n 25 gpurun_out/pytest_gpu2.log
