#!/bin/bash
# First-contact probe on the GPU box: every test in its own process under a timeout so a
# hung kernel cannot take the rest of the call with it.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for t in "$@"; do
  echo "=== $t" | tee -a gpurun_out/probe.log
  timeout 240 python -m pytest "tests/test_gpu_parity.py::$t" -x -q -m gpu 2>&1 | tail -25 | tee -a gpurun_out/probe.log
done
