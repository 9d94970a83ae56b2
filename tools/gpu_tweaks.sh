#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tweaks.log
timeout 300 python -m pytest tests -q -m gpu -x -k "stem or uint8 or posembed" > gpurun_out/stem_tests.log 2>&1
tail -4 gpurun_out/stem_tests.log
for v in "X=1" "LMV_POS_WAVES=1" "LMV_POS_WAVES=3"; do
  echo "== $v" | tee -a gpurun_out/tweaks.log
  env $v timeout 200 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1 2>&1 | grep -E "^posln|^stem|^\{\"model" | tee -a gpurun_out/tweaks.log
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['gpu_launches'])" | tee -a gpurun_out/tweaks.log
