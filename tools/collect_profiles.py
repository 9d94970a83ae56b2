"""Copy the evidence of a GPU round from gpurun_out/ (scratch) into profiles/ (tracked): bench lines, ncu launch list + per-kernel summary,
text briefs of the full captures, per-op CUDA-event profile, test logs, cycle traces, SASS mnemonics.
usage: python tools/collect_profiles.py <bench tag> <ncu tag> [out tag]      e.g.  r02c r02b r02"""
import glob, gzip, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
btag, ntag = sys.argv[1], sys.argv[2]
out = sys.argv[3] if len(sys.argv) > 3 else "r02"


def last_json(path):
    for line in reversed(open(path).read().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise SystemExit(f"no JSON line in {path}")


def copy(src, dst):
    if os.path.exists(src):
        shutil.copyfile(src, os.path.join(P, dst))
        print("copied", dst)
    else:
        print("MISSING", src)


for w, name in (("", "base256"), ("_tiny256", "tiny256"), ("_small512", "small512"), ("_base512seg", "base512seg")):
    json.dump(last_json(os.path.join(G, f"bench_{btag}{w}.log")), open(os.path.join(P, f"{out}_bench_{name}.json"), "w"))
json.dump(last_json(os.path.join(G, f"bench_ref_{ntag}.log")), open(os.path.join(P, f"{out}_bench_reference_arm.json"), "w"))
with open(os.path.join(G, f"launches_{ntag}.csv"), "rb") as f, gzip.open(os.path.join(P, f"{out}_launches_base_b256.csv.gz"), "wb") as g:
    g.write(f.read())
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summarize.py"), "launches", os.path.join(G, f"launches_{ntag}.csv"),
                os.path.join(P, f"{out}_kernels.json")], check=True)
d = json.load(open(os.path.join(P, f"{out}_kernels.json")))
d["source"] = f"profiles/{out}_launches_base_b256.csv.gz"
json.dump(d, open(os.path.join(P, f"{out}_kernels.json"), "w"), indent=1)
for rep in sorted(glob.glob(os.path.join(G, f"{ntag}_*.ncu-rep"))):
    name = os.path.basename(rep)[len(ntag) + 1:-len(".ncu-rep")]
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_brief.py"), rep, "18"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{out}_ncu_{name}.txt"), "w").write(txt)
    print("brief", name)
copy(os.path.join(G, f"ops_{btag}.log"), f"{out}_ops_base_b256_events.txt")
copy(os.path.join(G, f"ops_small_{btag}.log"), f"{out}_ops_small_b512_events.txt")
copy(os.path.join(G, f"smi_{ntag}.txt"), f"{out}_smi.txt")
copy(os.path.join(G, f"dca_trace_{ntag}.txt"), f"{out}_dca_role_trace.txt")
copy(os.path.join(G, f"attn_trace_{ntag}.txt"), f"{out}_attn_phase_trace.txt")
copy(os.path.join(G, f"mlp_trace_{ntag}.txt"), f"{out}_mlp_pair_trace.txt")
copy(os.path.join(G, f"mlp_ab_{ntag}.txt"), f"{out}_mlp_schedule_ab.txt")
with open(os.path.join(P, f"{out}_pytest_gpu.txt"), "w") as f:
    f.write(open(os.path.join(G, f"pytest_gpu_{btag}.log")).read()[-600:])
    f.write(open(os.path.join(G, f"smoke_{btag}.log")).read()[-300:])
open(os.path.join(P, f"{out}_sass_mnemonics.txt"), "w").write(
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_mnemonics.py")], capture_output=True, text=True).stdout)
