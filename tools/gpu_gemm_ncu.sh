#!/bin/bash
# tests of the current build + full ncu captures (source-level) of the K=384 GEMMs of the 'S' blocks: fc1 (LN fold + GELU), qkv (LN fold), proj (residual + statistics)
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-300 | tee -a gpurun_out/summary.txt; }
run gpu_tests 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q --maxfail=12
TAILN=100 run ops_base256 300 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1
CUTW=5000 TAILN=3 run bench_quick 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference
M=gpu__time_duration.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,dram__bytes_read.sum,dram__bytes_write.sum
cap() { run ncu_$1 400 ncu --set full --metrics $M --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/r02_$1 python tools/ncu_target.py lemevit_base 256 1; }
cap gemm_s_fc1 'tcgen05<3>' 2
cap gemm_s_qkv 'tcgen05<1>' 2
cap gemm_s_proj 'tcgen05<12>' 2
