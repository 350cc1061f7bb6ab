#!/bin/bash
# last check of the round: the GPU suite and the smoke test at the final tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_final.log
tail -n 3 gpurun_out/pytest_gpu_final.log gpurun_out/smoke_final.log
