#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/e2e.log
for rep in 1 2; do
for w in base256 tiny256 small512 base512seg; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(d['value']), round(d['e2e']['value']))" | tee -a gpurun_out/e2e.log
done
done
