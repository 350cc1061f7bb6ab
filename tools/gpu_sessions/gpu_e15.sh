#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-configs --lattice 6 --robots 1024 --selfcol --sparse 0.55 $2 > gpurun_out/b.json 2> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('$1 $2 value %.4e e2e %.4e strides %s batches %s'%(d['value'], d['e2e']['value'], d['config']['kernel']['strides'], d['config']['kernel']['batches']))" || tail -3 gpurun_out/b.err; }
run "fine"
VXG_STRIDES=224 run "one-stride-224"
VXG_STRIDES=128,224 run "two-strides"
VXG_STRIDES=192 run "one-stride-192"
run "fine-chunk1024" "--chunk-robots 1024"
VXG_STRIDES=224 run "224-chunk1024" "--chunk-robots 1024"
VXG_KERNEL_FORM=push VXG_STRIDES=128,224 run "push-two-strides"
# no self-collision, same sparse robots
timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-configs --lattice 6 --robots 1024 --sparse 0.55 > gpurun_out/b.json 2> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('sparse nocol fine value %.4e e2e %.4e'%(d['value'], d['e2e']['value']))"
VXG_STRIDES=224 timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-configs --lattice 6 --robots 1024 --sparse 0.55 > gpurun_out/b.json 2> gpurun_out/b.err; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('sparse nocol 224 value %.4e e2e %.4e'%(d['value'], d['e2e']['value']))"
