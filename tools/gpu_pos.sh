#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "posembed or model or golden or backbone or large" > gpurun_out/pos_tests.log 2>&1
tail -4 gpurun_out/pos_tests.log
timeout 200 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1 2>&1 | grep -E "^posln|^\{\"model|^meta_down" | tee gpurun_out/pos_ops.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['classes']['posembed_layernorm'])" | tee -a gpurun_out/pos_ops.log
