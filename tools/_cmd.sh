timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
LEMEVIT_B200_LANES=1 timeout 150 python tools/quick_bench.py lemevit_base 256 --ops 2>&1 | grep -E "lemevit_b200\]|rror|^posln|^gemm M=54272 N=1152" | cut -c1-130
timeout 100 python tools/quick_bench.py lemevit_base 256 --graph | head -1 | cut -c1-140
