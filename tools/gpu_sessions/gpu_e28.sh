#!/bin/bash
mkdir -p gpurun_out
VXG_TIMING2=1 timeout 900 python tools/ga_generation.py --robots 1024 --resident 5 --ref-sample 0 > gpurun_out/ga_e28.log 2>&1; grep -E "voxelyze:" gpurun_out/ga_e28.log | cut -c100-300
df /tmp | tail -1; mount | grep -E " /tmp | / " | head -3
