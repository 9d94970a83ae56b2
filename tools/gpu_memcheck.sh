#!/bin/bash
# compute-sanitizer memcheck: tools/gpu_memcheck.sh "<pytest -k expression>" [tag]
mkdir -p gpurun_out
timeout ${MEMCHECK_TIMEOUT:-200} compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests -q -m gpu -x -k "$1" > gpurun_out/memcheck_${2:-r02}.log 2>&1
echo "exit $?"
tail -8 gpurun_out/memcheck_${2:-r02}.log
