#!/bin/bash
# same-box A/B of two builds of the library: the shipped one and lemevit_b200/liblemevit_b200_prev.so
mkdir -p gpurun_out
rm -f gpurun_out/ab_lib.log
timeout 900 python -m pytest tests -q -m gpu --maxfail=5 > gpurun_out/gpu_tests.log 2>&1
tail -3 gpurun_out/gpu_tests.log
for v in "X=1" "LEMEVIT_B200_LIB=$PWD/lemevit_b200/liblemevit_b200_prev.so"; do
  echo "== $v" | tee -a gpurun_out/ab_lib.log
  env $v timeout 200 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1 2>&1 | grep -E "^gemm|^\{\"model" >> gpurun_out/ab_lib.log
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['classes']['gemm_tcgen05'])" | tee -a gpurun_out/ab_lib.log
done
