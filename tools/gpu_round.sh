#!/bin/bash
# One gpurun call at the end of a work phase: parity tests, smoke, bench (both arms), per-op profile, ncu launch list (with DRAM
# traffic) of the bench command + a full capture of the dominant kernel.  Every stage in its own process under a timeout.
# usage: tools/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
run() { # name, timeout, cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n ${TAILN:-12} gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
[ -z "$SKIP_TESTS" ] && run pytest_gpu 400 python -m pytest tests -m gpu -x -q
run smoke 200 python -c "import __graft_entry__ as g; g.smoke()"
run bench 400 python bench.py --steps 20 --warmup 5
[ -z "$SKIP_REF" ] && run bench_ref 300 python bench.py --impl reference --steps 3 --warmup 1
TAILN=80 run ops_base256 200 python tools/quick_bench.py lemevit_base 256 --ops
if [ -z "$SKIP_NCU" ]; then
  run ncu_launches 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-graph --no-e2e
  run ncu_full 400 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-gemm_bf16} -s ${NCU_SKIP:-120} -c 3 -f \
      -o gpurun_out/prof_$TAG python tools/ncu_target.py lemevit_base 256 2
fi
