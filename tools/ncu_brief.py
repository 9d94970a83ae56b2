"""Brief text summary of a --set full ncu report: headline counters + the source lines with the most stall samples.
usage: python tools/ncu_brief.py report.ncu-rep [n_lines]"""
import csv, subprocess, sys, collections, io
path = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]
for w in want:
    for i, h in enumerate(hdr):
        if h == w or h.endswith(w) and (w.startswith("launch__occupancy") ):
            print(f"{h:80s} {units[i]:10s} {vals[i][:90]}")
        elif w == "launch__occupancy_limit" and h.startswith(w):
            print(f"{h:80s} {units[i]:10s} {vals[i][:90]}")
# stall reasons (warp state sampling)
st = {h: vals[i] for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
tot = sum(float(v.replace(",", "") or 0) for v in st.values()) or 1
print("stall samples:", ", ".join(f"{k.replace('smsp__pcsamp_warps_issue_stalled_', '')} {float(v.replace(',', '') or 0) / tot * 100:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -float(kv[1].replace(',', '') or 0))[:8]))
src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rd = list(csv.reader(io.StringIO(src)))
hi = next((i for i, r in enumerate(rd) if r and r[0] == "Line No"), None)
if hi is not None:
    h = rd[hi]
    isamp = h.index("# Samples")
    iexec = h.index("Instructions Executed")
    stall_cols = [(i, x) for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    agg = collections.OrderedDict()
    cur = None
    for r in rd[hi + 1:]:
        if len(r) <= isamp:
            continue
        if r[0].strip():
            cur = (r[0], r[1].strip()[:130])
        if cur is None:
            continue
        a = agg.setdefault(cur, {"s": 0.0, "n": 0.0, "st": collections.Counter()})
        try:
            a["s"] += float(r[isamp] or 0)
            a["n"] += float(r[iexec] or 0)
            for i, x in stall_cols:
                a["st"][x] += float(r[i] or 0)
        except ValueError:
            pass
    tots = sum(a["s"] for a in agg.values()) or 1
    print(f"--- top source lines by stall samples (total {tots:.0f}, instructions {sum(a['n'] for a in agg.values()):.0f})")
    for (ln, text), a in sorted(agg.items(), key=lambda kv: -kv[1]["s"])[:nl]:
        top = ",".join(f"{k[6:]}:{v / max(a['s'], 1) * 100:.0f}" for k, v in a["st"].most_common(2))
        print(f"{a['s'] / tots * 100:5.1f}% L{ln:>4s} [{top:28s}] {text}")
