export LEMEVIT_B200_LIB=$PWD/lemevit_b200/liblemevit_b200_trace.so
mkdir -p gpurun_out
timeout 120 python tools/attn_trace.py 256 12 212 196 > gpurun_out/attn_events.txt 2>&1
