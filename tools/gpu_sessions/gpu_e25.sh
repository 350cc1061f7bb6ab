#!/bin/bash
# full GPU suite + smoke + both bench arms at the current tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r02f.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02f.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r02f.log
python bench.py --impl reference > gpurun_out/bench_ref_r02f.json 2> gpurun_out/bench_ref_r02f.err
python bench.py > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err
tail -n 3 gpurun_out/pytest_gpu_r02f.log gpurun_out/smoke_r02f.log; cat gpurun_out/bench_r02f.json | cut -c1-3000; cat gpurun_out/bench_ref_r02f.json | cut -c1-600; tail -n 3 gpurun_out/bench_r02f.err
