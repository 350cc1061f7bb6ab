#!/bin/bash
# One GPU session producing what profiles/ holds: tests, smoke, both bench arms, the ncu launch list of the bench
# command and one full ncu capture of the step kernel at the bench shape.   usage (under gpurun): tools/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log
python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vx_step_kernel -s 4 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -n 3 gpurun_out/pytest_gpu_$TAG.log gpurun_out/smoke_$TAG.log; cat gpurun_out/bench_$TAG.json gpurun_out/bench_ref_$TAG.json; tail -n 3 gpurun_out/bench_$TAG.err gpurun_out/ncu_full_$TAG.log
