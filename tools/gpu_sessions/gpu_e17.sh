#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_e17_full.log 2>&1; tail -6 gpurun_out/pytest_e17_full.log | cut -c1-250
run() { timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-configs $2 > gpurun_out/b.json 2> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('$1 | $2 | value %.4e frac %.3f e2e %.4e strides %s batches %s'%(d['value'], d['roofline']['frac'], d['e2e']['value'], d['config']['kernel']['strides'], d['config']['kernel']['batches']))" || tail -3 gpurun_out/b.err; }
run rt "--lattice 6 --robots 1024 --selfcol --sparse 0.55"
VXG_RT_STRIDE=0 run ct "--lattice 6 --robots 1024 --selfcol --sparse 0.55"
run rt "--lattice 6 --robots 1024 --sparse 0.55"
VXG_RT_STRIDE=0 run ct "--lattice 6 --robots 1024 --sparse 0.55"
run rt "--lattice 6 --robots 4096"
VXG_RT_STRIDE=0 run ct "--lattice 6 --robots 4096"
run rt "--lattice 6 --robots 4096 --selfcol"
VXG_RT_STRIDE=0 run ct "--lattice 6 --robots 4096 --selfcol"
run rt "--lattice 4 --robots 8192"
VXG_RT_STRIDE=0 run ct "--lattice 4 --robots 8192"
