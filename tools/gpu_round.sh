#!/bin/bash
# one GPU session: tests, smoke, bench (both arms), ncu launch list + full capture of the step kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --timesteps 100 --no-cpu-baseline > gpurun_out/bench_ts100.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vx_step_kernel -s 2 -c 1 -f -o gpurun_out/prof_step \
    python bench.py --steps 3 --warmup 3 --timesteps 100 --robots 296 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log; cat gpurun_out/bench.json gpurun_out/bench_ref.json gpurun_out/bench_ts100.json; tail -5 gpurun_out/bench.err gpurun_out/ncu_full.log
