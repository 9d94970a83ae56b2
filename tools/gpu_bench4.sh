#!/bin/bash
# tests + the four workload bench lines (both reference arms inside bench.py) — the final numbers of a round: tools/gpu_bench4.sh [tag]
TAG=${1:-r02}
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a gpurun_out/summary.txt; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-${CUTW:-300} | tee -a gpurun_out/summary.txt; }
run pytest_gpu_$TAG 1200 python -m pytest tests -m gpu -q
run smoke_$TAG 300 python -c "import __graft_entry__ as g; g.smoke()"
CUTW=8000 TAILN=2 run bench_$TAG 700 python bench.py --steps 20 --warmup 5
for w in tiny256 small512 base512seg; do CUTW=8000 TAILN=2 run bench_${TAG}_$w 600 python bench.py --workload $w --steps 20 --warmup 5; done
TAILN=100 run ops_$TAG 300 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1
TAILN=100 run ops_small_$TAG 300 python tools/quick_bench.py lemevit_small 512 --ops --lanes=1
