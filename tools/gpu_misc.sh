mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --maxfail=10 > gpurun_out/gpu_tests.log 2>&1
timeout 200 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1 > gpurun_out/ops_base256.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference > gpurun_out/bench_q.log 2>&1
python bench.py --workload base512seg --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference > gpurun_out/bench_q512.log 2>&1
