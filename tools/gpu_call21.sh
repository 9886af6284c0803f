#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/k1_points.py 2>&1 | tail -8
timeout 900 python tools/ncu_traffic.py > gpurun_out/ncu_traffic.log 2>&1; echo "ncu_traffic rc=$?"; tail -5 gpurun_out/ncu_traffic.log
