#!/bin/bash
# last check of a round: the commands the driver runs (GPU tests, smoke, default bench), on the shipped library
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/verify_tests.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/verify_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/verify_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/verify_smoke.log
timeout 600 python bench.py > gpurun_out/verify_bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/verify_bench.log | cut -c1-400
