#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches_prefill_c4.csv python bench.py --config 4 --steps 1 --warmup 1 --new-tokens 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_prefill_c4.log 2>&1
tail -n 2 gpurun_out/ncu_prefill_c4.log | cut -c1-200
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "prefill_natural or teacher_forced" > gpurun_out/pytest_tol.log 2>&1; tail -n 5 gpurun_out/pytest_tol.log
