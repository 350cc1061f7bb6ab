#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "sticky or all_reference_example or import" > gpurun_out/pytest_e26.log 2>&1; tail -6 gpurun_out/pytest_e26.log | cut -c1-400
