#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/robustness_smoke.py 2>&1 | tail -25
