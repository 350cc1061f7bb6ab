#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_e3_full.log 2>&1; tail -80 gpurun_out/pytest_e3_full.log | cut -c1-300
