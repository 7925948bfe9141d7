#!/bin/bash
# runs every probe step in its own process with a timeout; log goes to gpurun_out/probe.log
mkdir -p gpurun_out
for s in "$@"; do
  echo "=== $s ===" | tee -a gpurun_out/probe.log
  timeout 300 python tools/gpu_probe.py $s 2>&1 | tail -40 | tee -a gpurun_out/probe.log
  echo "exit=$?" | tee -a gpurun_out/probe.log
done
