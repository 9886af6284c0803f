#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log | cut -c1-300
for i in 1 2 3 4 5 6; do
  timeout 300 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -m gpu -x -q -k "full_size or repeatable or determinism or graph or deadlock" > gpurun_out/flake_$i.log 2>&1
  echo "flake run $i rc=$? $(tail -n 1 gpurun_out/flake_$i.log)"
done
