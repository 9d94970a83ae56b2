#!/bin/bash
# DCA bring-up: kernel tests with the barrier-tag debug build first, then the release build, then model tests + profile + bench
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-25} gpurun_out/$name.log | cut -c1-${CUTW:-300} | tee -a gpurun_out/summary.txt; }
LEMEVIT_B200_LIB=$PWD/lemevit_b200/liblemevit_b200_dbg.so run dca_dbg_small 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "dca_block_fused and (3-300-96-3 or 2-784-192-6 or 3-128-32-1)" -s
run dca_kernels 900 python -m pytest tests/test_gpu_kernels.py -q -k "dca" --maxfail=30 -s
run gpu_model 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q --maxfail=12
TAILN=100 run ops_base256 300 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1
CUTW=5000 TAILN=3 run bench_quick 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference
export LEMEVIT_B200_LIB=$PWD/lemevit_b200/liblemevit_b200_trace.so
for a in "128 3136 96 3 D 0" "128 3136 96 3 C 0" "128 784 192 6 D 0"; do timeout 120 python tools/dca_trace.py $a; done 2>&1 | tee gpurun_out/dca_trace.txt
