#!/bin/bash
mkdir -p gpurun_out
timeout 700 python tools/pdl_bitwise.py gpurun_out/pdl_on.pt 2>&1 | tail -2
DLB200_NO_PDL=15 timeout 700 python tools/pdl_bitwise.py gpurun_out/pdl_off.pt 2>&1 | tail -2
python tools/pdl_bitwise.py --compare gpurun_out/pdl_on.pt gpurun_out/pdl_off.pt 2>&1 | tail -20
rm -f gpurun_out/pdl_on.pt gpurun_out/pdl_off.pt
