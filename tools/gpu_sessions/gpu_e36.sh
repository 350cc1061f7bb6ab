#!/bin/bash
mkdir -p gpurun_out
for C in 128 64 256 1024; do
VXG_CHUNK_ROBOTS=$C timeout 900 python tools/ga_generation.py --robots 1024 --resident 3 --ref-sample 0 > gpurun_out/ga_e36.log 2>&1; echo "chunk $C"; grep -E "voxelyze:" gpurun_out/ga_e36.log | cut -c100-200 | tail -2
done
