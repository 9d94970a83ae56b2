#!/bin/bash
# images-per-CTA sweep of the meta-token kernels: tests, then per-op profile and bench line for LMV_META_IM = 1 / 2 / 4
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --maxfail=10 > gpurun_out/gpu_tests.log 2>&1
tail -5 gpurun_out/gpu_tests.log
for im in 1 2 4; do
  LMV_META_IM=$im timeout 200 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1 2>&1 | grep -E "^meta|^\{\"model" > gpurun_out/meta_im$im.log
  LMV_META_IM=$im timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('IM=$im', d['value'], d['e2e']['value'], d['roofline']['classes']['meta_branch'])" | tee -a gpurun_out/meta_im$im.log
done
