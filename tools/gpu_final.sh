#!/bin/bash
# Final evidence round: tests, smoke, the four workloads with both reference arms, per-op profile, ncu launch list with tensor metrics,
# full ncu captures of every kernel class, cycle traces.  usage: tools/gpu_final.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi_$TAG.txt 2>&1
nproc >> gpurun_out/smi_$TAG.txt
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-${CUTW:-300} | tee -a gpurun_out/summary.txt; }
run pytest_gpu_$TAG 1200 python -m pytest tests -m gpu -q
run smoke_$TAG 300 python -c "import __graft_entry__ as g; g.smoke()"
CUTW=8000 TAILN=2 run bench_$TAG 700 python bench.py --steps 20 --warmup 5
for w in tiny256 small512 base512seg; do CUTW=8000 TAILN=2 run bench_${TAG}_$w 600 python bench.py --workload $w --steps 20 --warmup 5; done
CUTW=3000 TAILN=2 run bench_ref_$TAG 500 python bench.py --impl reference --steps 3 --warmup 1
TAILN=100 run ops_$TAG 300 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1
M=gpu__time_duration.sum,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
run ncu_launches_$TAG 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tools/ncu_target.py lemevit_base 256 2
cap() { run ncu_$1 400 ncu --set full --metrics $M --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/${TAG}_$1 python tools/ncu_target.py lemevit_base 256 1; }
cap dca_d_c96 dca_x_kernel 2
cap dca_d_c192 dca_x_kernel 6
cap dca_c_c96 dca_x_kernel 0
cap gemm_qkv gemm_bf16 ${GEMM_SKIP:-30}
cap mlp_c96 mlp_fused_tcgen05 0
cap mlp_pair_c384 mlp_pair 2
cap attn_self attention_self_kernel 2
cap posembed_c96 posembed_tile 0
cap posembed_c384 posembed_tile 22
cap meta_chain_c192 meta_chain 8
# cycle traces (debug builds, if present): fused cross-attention roles, self-attention softmax phases, CTA-pair MLP issuer / epilogue
if [ -f lemevit_b200/liblemevit_b200_dcatrace.so ]; then
  export LEMEVIT_B200_LIB=$PWD/lemevit_b200/liblemevit_b200_dcatrace.so
  for a in "128 3136 96 3 D 0" "128 3136 96 3 C 0" "128 784 192 6 D 0"; do timeout 120 python tools/dca_trace.py $a; done > gpurun_out/dca_trace_$TAG.txt 2>&1
fi
if [ -f lemevit_b200/liblemevit_b200_attntrace.so ]; then
  export LEMEVIT_B200_LIB=$PWD/lemevit_b200/liblemevit_b200_attntrace.so
  timeout 120 python tools/attn_trace.py 256 12 212 196 > gpurun_out/attn_trace_$TAG.txt 2>&1
fi
if [ -f lemevit_b200/liblemevit_b200_mlptrace.so ]; then
  export LEMEVIT_B200_LIB=$PWD/lemevit_b200/liblemevit_b200_mlptrace.so
  timeout 120 python tools/mlp_trace.py > gpurun_out/mlp_trace_$TAG.txt 2>&1
fi
unset LEMEVIT_B200_LIB
# schedule A/B lines behind the numbers in DESIGN.md: stage-3 MLP as two GEMMs / single-CTA wide kernel / CTA-pair kernel
for v in "LMV_FUSED_MLP_WIDE=0" "LMV_MLP_PAIR=0" "LMV_MLP_PAIR=1"; do echo "== $v"; env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference | tail -1 | cut -c1-200; done > gpurun_out/mlp_ab_$TAG.txt 2>&1
true
