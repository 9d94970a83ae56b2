#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/lanes2.log
for w in tiny256 small512 base512seg; do
  for l in 1 2 4; do
    LEMEVIT_B200_LANES=$l timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w lanes=$l', round(d['value']), round(d['e2e']['value']), d['config'].get('lanes'))" | tee -a gpurun_out/lanes2.log
  done
done
