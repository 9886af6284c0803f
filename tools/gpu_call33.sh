#!/bin/bash
mkdir -p gpurun_out
timeout 1900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_all.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
