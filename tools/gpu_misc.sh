mkdir -p gpurun_out
timeout 100 ./tools/micro/hbm_write_bench > gpurun_out/hbm_write.txt 2>&1
for c in 0 128 64; do for l in 1 2; do echo "== chunk $c lanes $l"; timeout 200 python tools/quick_bench.py lemevit_base 256 $c --graph --lanes=$l 2>&1 | tail -1 | cut -c1-200; done; done > gpurun_out/chunk_sweep.txt 2>&1
