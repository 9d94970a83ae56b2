#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-15} gpurun_out/$name.log | cut -c1-300 | tee -a gpurun_out/summary.txt; }
run self_odd 300 python -m pytest tests/test_gpu_kernels.py -q -k "attention_self_two_segments" --maxfail=3
run large 600 python -m pytest tests/test_gpu_large.py -q -s
run all_gpu 1200 python -m pytest tests -m gpu -q --maxfail=10
