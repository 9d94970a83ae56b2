#!/bin/bash
# regression check: full GPU test suite, per-op profile (one lane), headline bench
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-300 | tee -a gpurun_out/summary.txt; }
run gpu_tests 1500 python -m pytest tests -q -m gpu --maxfail=12
TAILN=100 run ops_base256 300 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1
CUTW=5000 TAILN=3 run bench_quick 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference
${EXTRA_CMD:-true}
