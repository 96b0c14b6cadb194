#!/bin/bash
# Runs each tc_probe test in its own process (a trap in one test must not poison the next).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/probe_gpu.txt 2>&1
for t in "$@"; do
  echo "=== $t ===" | tee -a gpurun_out/probe.log
  timeout 120 ./build/tc_probe $t 2>&1 | tee -a gpurun_out/probe.log
  echo "exit: ${PIPESTATUS[0]}" | tee -a gpurun_out/probe.log
done
