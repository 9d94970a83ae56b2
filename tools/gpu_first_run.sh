#!/bin/bash
# Bring-up driver for a gpurun call: every stage in its own process (a trapped kernel poisons only that stage),
# every stage under a timeout, all logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
"$@"
