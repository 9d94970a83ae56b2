#!/bin/bash
# CTA-pair MLP bring-up: small kernel tests first (bounded), then the rest
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-300 | tee -a gpurun_out/summary.txt; }
run pair_small 120 python -m pytest tests/test_gpu_kernels.py -q -x -k "mlp_fused and 384" || exit 0
grep -q "passed" gpurun_out/pair_small.log || exit 0
grep -q "failed" gpurun_out/pair_small.log && exit 0
run gpu_tests 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q --maxfail=12
TAILN=100 run ops_base256 300 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1
CUTW=5000 TAILN=3 run bench_quick 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference
LMV_MLP_PAIR=0 CUTW=5000 TAILN=3 run bench_quick_nopair 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference
