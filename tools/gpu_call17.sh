#!/bin/bash
# float16 build: fp16 fixture tests + the GPU suites re-run in fp16, one full-size fp16 parity row, an fp16 bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_fp16_gpu.py -x -q > gpurun_out/fp16_tests.log 2>&1; echo "fp16 tests rc=$?" | tee -a gpurun_out/fp16_tests.log
tail -30 gpurun_out/fp16_tests.log
DLB200_PARITY_REPORT=gpurun_out/fp16_parity.jsonl timeout 600 python -m pytest tests/test_full_size_parity_gpu.py -q -k float16 > gpurun_out/fp16_full.log 2>&1; echo "fp16 full-size rc=$?" | tee -a gpurun_out/fp16_full.log
tail -15 gpurun_out/fp16_full.log
timeout 600 python bench.py --dtype fp16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; echo "bench fp16 rc=$?"
tail -c 1500 gpurun_out/bench_fp16.json
