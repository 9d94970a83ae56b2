#!/bin/bash
# full ncu captures (source-level) of the stage-3 kernels after the round's changes + regression tests
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-300 | tee -a gpurun_out/summary.txt; }
run gpu_tests 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q --maxfail=12
TAILN=100 run ops_base256 300 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1
M=gpu__time_duration.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,dram__bytes_read.sum,dram__bytes_write.sum
cap() { run ncu_$1 400 ncu --set full --metrics $M --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/r02_$1 python tools/ncu_target.py lemevit_base 256 1; }
cap posembed_c384 posembed_tile 22
cap attn_self attention_self_kernel 2
cap mlp_pair mlp_pair 2
