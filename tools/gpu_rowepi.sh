#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/rowepi.log
timeout 900 python -m pytest tests -q -m gpu --maxfail=5 > gpurun_out/gpu_tests.log 2>&1
tail -6 gpurun_out/gpu_tests.log
for v in "X=1" "LMV_GEMM_ROW_RES=0"; do
  echo "== $v" | tee -a gpurun_out/rowepi.log
  env $v timeout 200 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1 2>&1 | grep -E "^gemm|^\{\"model" >> gpurun_out/rowepi.log
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['classes']['gemm_tcgen05'])" | tee -a gpurun_out/rowepi.log
done
