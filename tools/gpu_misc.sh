mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q -x -k "mlp" > gpurun_out/mlp_tests.log 2>&1
LEMEVIT_B200_LIB=$PWD/lemevit_b200/liblemevit_b200_mlptrace.so timeout 120 python tools/mlp_trace.py --debug=0 > gpurun_out/mlp_trace.txt 2>&1
timeout 200 python tools/quick_bench.py lemevit_base 256 --ops --lanes=1 > gpurun_out/ops_base256.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-reference > gpurun_out/bench_q.log 2>&1
