#!/bin/bash
# Round-2 gpurun script: parity tests, smoke, the four bench workloads (with the eager-GPU reference arm and the CPU arm), the
# per-op profile, and ncu passes.  Every stage in its own process under a timeout.  Stages are selected with STAGES="a b c".
# usage: STAGES="tests bench workloads ncu_metrics" tools/gpu_round2.sh [tag]
TAG=${1:-r02}
STAGES=${STAGES:-tests smoke bench workloads ref ops}
mkdir -p gpurun_out
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
has() { [[ " $STAGES " == *" $1 "* ]]; }
run() { # name, timeout, cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-${CUTW:-400} | tee -a gpurun_out/summary.txt
}
NCU_METRICS=gpu__time_duration.sum,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
has tests && run pytest_gpu 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS}
has smoke && run smoke 200 python -c "import __graft_entry__ as g; g.smoke()"
has bench && CUTW=6000 run bench_$TAG 600 python bench.py --steps 20 --warmup 5
if has workloads; then
  for w in tiny256 small512 base512seg; do CUTW=6000 run bench_${TAG}_$w 500 python bench.py --workload $w --steps 20 --warmup 5; done
fi
has ref && CUTW=3000 run bench_ref_$TAG 400 python bench.py --impl reference --steps 3 --warmup 1
has ops && TAILN=90 run ops_base256 200 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1
if has ncu_metrics; then
  # per-launch device time, DRAM traffic and tensor-pipe work (UTCHMMA math ops: a per-launch counter, unlike the *_realtime ones)
  run ncu_launches 900 ncu --metrics $NCU_METRICS --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
      python tools/ncu_target.py lemevit_base 256 2
fi
if has ncu_full; then
  for k in ${NCU_KERNELS:-gemm_bf16}; do
    run ncu_full_$k 500 ncu --set full --metrics $NCU_METRICS --clock-control none --import-source on -k regex:$k -s ${NCU_SKIP:-20} -c ${NCU_COUNT:-3} -f \
        -o gpurun_out/prof_${TAG}_$k python tools/ncu_target.py lemevit_base 256 1
  done
fi
true
