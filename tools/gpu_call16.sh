#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck_model.log python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "not full_size" > gpurun_out/memcheck_model_pytest.log 2>&1
echo "memcheck(model) rc=$?"; tail -n 2 gpurun_out/memcheck_model_pytest.log; tail -n 2 gpurun_out/memcheck_model.log
rm -f gpurun_out/full_size_parity.jsonl
timeout 300 python - <<'PY' > gpurun_out/parity_seed1.log 2>&1
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import test_full_size_parity_gpu as T
from dynamic_llava_b200.synthetic import forced_text_decisions
fd = forced_text_decisions(63, 8, keep_rate=0.5, seed=5)
T._run_config("2 (ragged default, weight seed 1, 64 steps)", "llava-1.5-7b", 8, 32, 64, 0.2, fd, quirk=False, per_sample_rows=[2, 5], seed=1)
fd = forced_text_decisions(63, 8, keep_rate=0.5, seed=5)
T._run_config("2 (reference batched semantics, weight seed 2, 64 steps)", "llava-1.5-7b", 8, 32, 64, 0.2, fd, quirk=True, seed=2)
PY
tail -n 3 gpurun_out/parity_seed1.log | cut -c1-400
