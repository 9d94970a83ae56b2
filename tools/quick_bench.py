"""Quick device-time probe (not the contract bench): CUDA-event timing of the native forward and a
torch.profiler kernel table.  usage: python tools/quick_bench.py lemevit_base 256 [chunk] [--graph] [--prof]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lemevit_b200 as L
from oracle import lemevit_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "lemevit_tiny"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
chunk = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 0
use_graph = "--graph" in sys.argv
prof = "--prof" in sys.argv
res = 224
torch.manual_seed(0)
m = getattr(L, name)(native_chunk=chunk).to("cuda", torch.bfloat16)
m.train(False)
x = torch.randn(B, 3, res, res, device="cuda").to(torch.bfloat16)
with torch.no_grad():
    for a in sys.argv:
        if a.startswith("--opt="):          # e.g. --opt=fused_mlp=0
            k, v = a[6:].split("=")
            m.native_engine(x.device).set_option(k, int(v))
    eng = m.native_engine(x.device)
    for a in sys.argv:
        if a.startswith("--lanes="):        # --lanes=1: kernels of a launch bracket run alone (per-op profile)
            eng.lanes = int(a[8:])
    for _ in range(3):
        y = m(x)
    torch.cuda.synchronize()
    fn = (lambda: m(x))
    if use_graph:
        sx, sy, replay = eng.graphed(x)
        fn = replay
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    iters = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.time() - t0) / iters * 1e3
    ms = e0.elapsed_time(e1) / iters
    gflop = O.algorithmic_flops_per_image(O.VARIANTS[name], res, res) / 1e9
    print(json.dumps({"model": name, "batch": B, "chunk": chunk, "graph": use_graph, "ms": ms, "wall_ms": wall, "img_s": B / ms * 1e3,
                      "tflops": B * gflop / ms, "launches": eng.launch_count(B, res, res)}))
    if "--ops" in sys.argv:
        eng.set_profile(True)
        for _ in range(3):
            m(x)
        print(eng.profile_report())
        for e in eng.get_profile():
            print(e)
        eng.set_profile(False)
    if prof:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
            for _ in range(2):
                m(x)
            torch.cuda.synchronize()
        print(p.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
