#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/lanes.txt
for l in 1 2 3 4; do
  echo "lanes=$l" >> gpurun_out/lanes.txt
  LEMEVIT_B200_LANES=$l timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference --no-profile --no-e2e 2>&1 | grep -o '"value": [0-9.]*, "unit": "img/s", "n_gpus": 1, "steps": 20, "warmup": 5, "ms_per_step": [0-9.]*' >> gpurun_out/lanes.txt
done
cat gpurun_out/lanes.txt
