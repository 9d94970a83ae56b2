#!/bin/bash
# Evidence round without the debug-build cycle traces and the MLP schedule A/B (kernels unchanged since profiles/r02_*_trace.txt):
# tests, smoke, the four workloads with both reference arms, per-op profiles, ncu launch list, full captures.  usage: tools/gpu_final2.sh [tag]
TAG=${1:-r02d}
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
TAILN=100 run ops_small_$TAG 300 python tools/quick_bench.py lemevit_small 512 --ops --lanes=1
M=gpu__time_duration.sum,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
run ncu_launches_$TAG 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tools/ncu_target.py lemevit_base 256 2
cap() { run ncu_$1 400 ncu --set full --metrics $M --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/${TAG}_$1 python tools/ncu_target.py lemevit_base 256 1; }
cap gemm_qkv gemm_bf16 3
cap gemm_proj gemm_bf16 4
true
