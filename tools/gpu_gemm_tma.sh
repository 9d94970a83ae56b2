#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/gemm_tma_tests.log 2>&1
tail -4 gpurun_out/gemm_tma_tests.log
for v in 0 1; do
  echo "== LMV_GEMM_TMA_OUT=$v" | tee -a gpurun_out/gemm_tma_ops.log
  LMV_GEMM_TMA_OUT=$v timeout 200 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1 2>&1 | grep -E "^gemm|^\{\"model" | tee -a gpurun_out/gemm_tma_ops.log
  LMV_GEMM_TMA_OUT=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['classes']['gemm_tcgen05'])" | tee -a gpurun_out/gemm_tma_ops.log
done
