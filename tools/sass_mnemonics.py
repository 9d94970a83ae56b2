"""Per-kernel count of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTCHMMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA loads), UTCBAR (tcgen05.commit), HMMA (mma.sync), plus registers per thread.
usage: python tools/sass_mnemonics.py [lib.so] > profiles/rNN_sass_mnemonics.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lemevit_b200", "liblemevit_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
name = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        name = m.group(1)
    m = re.search(r"REG:(\d+)", line)
    if m and name:
        regs[name] = int(m.group(1))
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for mn in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "MUFU", "R2UR"):
            if re.search(r"\b" + mn + r"\b|\b" + mn + r"\.", line):
                counts[cur][mn] += 1
def demangle(n):
    try:
        d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
        return re.sub(r"^void ", "", d.split("(")[0]).replace("lmv::", "")
    except Exception:
        return n
print(f"# SASS mnemonics per kernel of {os.path.basename(lib)} (cuobjdump -sass); UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA load,")
print("# UTCBAR = tcgen05.commit, HMMA = warp-level mma.sync (meta-token kernels only), R2UR = register -> uniform-register moves")
print("# UTMASTG = TMA tile store (GEMM epilogue)")
print(f"{'kernel':58s} {'regs':>5s} {'UTCHMMA':>8s} {'LDTM':>6s} {'STTM':>6s} {'UTMALDG':>8s} {'UTMASTG':>8s} {'UTCBAR':>7s} {'HMMA':>6s} {'MUFU':>6s} {'R2UR':>6s}")
for k, c in counts.items():
    d = demangle(k)
    print(f"{d[:58]:58s} {regs.get(k, 0):5d} {c['UTCHMMA']:8d} {c['LDTM']:6d} {c['STTM']:6d} {c['UTMALDG']:8d} {c['UTMASTG']:8d} {c['UTCBAR']:7d} {c['HMMA']:6d} {c['MUFU']:6d} {c['R2UR']:6d}")
