#!/bin/bash
# Short gpurun call while iterating on a kernel: selected tests + per-op profile.  usage: tools/gpu_quick.sh "<pytest -k expr>" [quick_bench args...]
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n ${TAILN:-15} gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
K="$1"; shift
[ -n "$K" ] && run pytest_sel ${PYTEST_TIMEOUT:-240} python -m pytest tests -m gpu -x -q -k "$K"
if [ $# -gt 0 ]; then TAILN=90 run ops 200 python tools/quick_bench.py "$@"; fi
