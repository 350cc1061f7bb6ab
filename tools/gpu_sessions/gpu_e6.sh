#!/bin/bash
mkdir -p gpurun_out
VXG_KERNEL_FORM=barrier VXG_LIB=$PWD/evosoro_b200/lib_trace/libvxgpu.so timeout 300 python bench.py --no-cpu-baseline --steps 3 --robots 148 --timesteps 1000 2>&1 | grep TRACE | sort -k3,3n -k5,5n > gpurun_out/trace_e6.txt
head -70 gpurun_out/trace_e6.txt
