#!/bin/bash
# fused MLP iteration: kernel tests (debug wait tags first), model tests, per-op profile, bench, A/B against the unfused S-block MLP
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-300 | tee -a gpurun_out/summary.txt; }
run mlp_tests 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "mlp_fused"
run gpu_tests 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q --maxfail=12
TAILN=100 run ops_base256 300 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1
CUTW=5000 TAILN=3 run bench_quick 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference
LMV_MLP_WIDE=0 CUTW=5000 TAILN=3 run bench_quick_unfused 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference
${EXTRA_CMD:-true}
