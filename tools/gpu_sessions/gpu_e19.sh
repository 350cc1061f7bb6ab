#!/bin/bash
# cluster form (vxg_cluster.cuh): bit identity against the push form, then A/B: single-CTA gather vs cluster (release/acquire) vs cluster (relaxed arrive)
mkdir -p gpurun_out
for L in lib lib_v2; do
  VXG_LIB=$PWD/evosoro_b200/$L/libvxgpu.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gather_form_is_bit" > gpurun_out/pytest_e19_$L.log 2>&1; tail -3 gpurun_out/pytest_e19_$L.log | cut -c1-300
done
run() {  # label, lib, form, shape
  VXG_LIB=$PWD/evosoro_b200/$2/libvxgpu.so VXG_KERNEL_FORM=$3 timeout 300 python bench.py --no-cpu-baseline --no-configs --steps 5 $4 > gpurun_out/bench_e19.json 2> gpurun_out/bench_e19.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_e19.json"))
    print("$1 $4 value %.4e e2e %.4e frac %.3f ms %.2f clocks %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["ms_per_step"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$1 $4 FAILED", e); print(open("gpurun_out/bench_e19.err").read()[-600:])
PY
}
for SHAPE in "--lattice 10 --robots 1024" "--lattice 10 --robots 4096"; do
  run gather1 lib barrier "$SHAPE"
  run cluster_v1 lib cluster "$SHAPE"
  run cluster_v2 lib_v2 cluster "$SHAPE"
done 2>&1 | tee -a gpurun_out/exp_e19.txt
