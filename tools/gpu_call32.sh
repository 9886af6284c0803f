#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/graph_vs_eager.py llava-1.5-13b 2>&1 | tail -11
export PDL_BITWISE_PRESET=llava-1.5-13b
timeout 900 python tools/pdl_bitwise.py gpurun_out/pdl_on.pt 2>&1 | tail -1
DLB200_NO_PDL=15 timeout 900 python tools/pdl_bitwise.py gpurun_out/pdl_off.pt 2>&1 | tail -1
python tools/pdl_bitwise.py --compare gpurun_out/pdl_on.pt gpurun_out/pdl_off.pt 2>&1 | tail -8
rm -f gpurun_out/pdl_on.pt gpurun_out/pdl_off.pt
