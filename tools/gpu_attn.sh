#!/bin/bash
# self-attention kernel: tests, phase trace of the softmax warps, isolated timing
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-400 | tee -a gpurun_out/summary.txt; }
[ -z "$SKIP_TESTS" ] && run gpu_tests 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q --maxfail=12
export LEMEVIT_B200_LIB=$PWD/lemevit_b200/liblemevit_b200_trace.so
for a in "128 12 212 196" "256 12 212 196" "128 16 65 49" "8 12 1024 1024"; do timeout 120 python tools/attn_trace.py $a; done 2>&1 | tee gpurun_out/attn_trace.txt | cut -c1-400 | tee -a gpurun_out/summary.txt
unset LEMEVIT_B200_LIB
[ -n "$WITH_OPS" ] && TAILN=100 run ops_base256 300 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1
[ -n "$WITH_BENCH" ] && CUTW=5000 TAILN=3 run bench_quick 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference
true
