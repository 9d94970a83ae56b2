"""Condense ncu outputs into the small JSON / text summaries that are committed under profiles/.

  python tools/ncu_summarize.py launches <launches.csv> <out.json>     per-kernel shares + DRAM traffic from a metrics pass
        (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file launches.csv <cmd>)
  python tools/ncu_summarize.py full <prof.ncu-rep> <out.txt>          the headline counters of a --set full capture
"""
import collections
import csv
import json
import re
import subprocess
import sys


TENSOR_OPS = "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32.sum"   # per-launch FLOPs (2 x MAC, padded tiles included) issued by UTCHMMA


def short(name: str) -> str:
    name = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"^void\s+", "", name.strip())
    name = re.sub(r"\(.*$", "", name)          # argument list
    name = re.sub(r"<.*$", "", name)           # template arguments
    return name.split("::")[-1].strip()


def measured_peak_tflops():
    import os
    try:
        d = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        return float(d["bf16_tflops"]), "bf16_tflops (burst) of measured"
    except Exception:
        return 1590.0, "fallback"


def launches(path, out):
    rows = list(csv.DictReader(l for l in open(path, errors="replace") if l.startswith('"')))
    per_id = collections.OrderedDict()
    for r in rows:
        d = per_id.setdefault(r["ID"], {"kernel": short(r["Kernel Name"])})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r["Metric Unit"]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[r["Metric Name"]] = v * scale
    agg = collections.defaultdict(lambda: {"launches": 0, "time_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "tensor_flops": 0.0,
                                           "best_util": 0.0})
    peak, peak_src = measured_peak_tflops()
    for d in per_id.values():
        a = agg[d["kernel"]]
        a["launches"] += 1
        a["time_us"] += d.get("gpu__time_duration.sum", 0.0)
        a["dram_read_bytes"] += d.get("dram__bytes_read.sum", 0.0)
        a["dram_write_bytes"] += d.get("dram__bytes_write.sum", 0.0)
        a["tensor_flops"] += d.get(TENSOR_OPS, 0.0)
        t = d.get("gpu__time_duration.sum", 0.0)
        if t > 0:
            a["best_util"] = max(a["best_util"], d.get(TENSOR_OPS, 0.0) / (t * 1e-6) / 1e12 / peak)
    total = sum(a["time_us"] for a in agg.values()) or 1.0
    kernels = {}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["time_us"]):
        n = a["launches"]
        kernels[k] = {"launches": n, "time_us": round(a["time_us"], 1), "share": round(a["time_us"] / total, 4),
                      "avg_us": round(a["time_us"] / n, 2), "dram_bytes_per_launch": round((a["dram_read_bytes"] + a["dram_write_bytes"]) / n),
                      "dram_read_bytes_per_launch": round(a["dram_read_bytes"] / n), "dram_write_bytes_per_launch": round(a["dram_write_bytes"] / n),
                      "dram_gbs": round((a["dram_read_bytes"] + a["dram_write_bytes"]) / (a["time_us"] * 1e-6) / 1e9, 1) if a["time_us"] else 0.0,
                      # tensor-pipe utilisation of the launch = UTCHMMA FLOPs counted by the hardware / duration / measured cuBLAS bf16 peak
                      "tensor_tflops": round(a["tensor_flops"] / (a["time_us"] * 1e-6) / 1e12, 1) if a["time_us"] else 0.0,
                      "tensor_pipe_util": round(a["tensor_flops"] / (a["time_us"] * 1e-6) / 1e12 / peak, 4) if a["time_us"] else 0.0,
                      "tensor_pipe_util_best_launch": round(a["best_util"], 4)}
    json.dump({"source": path, "note": "ncu per-launch times are cold-cache and serialised: compare shares, not absolutes", "total_time_us": round(total, 1),
               "tensor_metric": TENSOR_OPS + " (per-launch counter; FLOPs incl. tile padding) / gpu__time_duration.sum / peak", "tensor_peak_tflops": peak,
               "tensor_peak_source": peak_src,
               "kernels": kernels}, open(out, "w"), indent=1)
    for k, v in list(kernels.items())[:12]:
        print(f"{k:34s} n={v['launches']:5d} share={v['share']*100:5.1f}% avg={v['avg_us']:8.1f}us dram/launch={v['dram_bytes_per_launch']/1e6:8.1f} MB "
              f"{v['dram_gbs']:7.1f} GB/s tensor {v['tensor_tflops']:7.1f} TFLOP/s = {v['tensor_pipe_util']*100:5.1f}% (best launch {v['tensor_pipe_util_best_launch']*100:5.1f}%)")


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_active.avg",
            "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {path} (one block per captured launch)\n")
        for r in rows[2:]:
            for w in want:
                key = next((h for h in hdr if h.endswith(w) and h in idx), None)
                if key:
                    f.write(f"{w:75s} {units[idx[key]]:14s} {r[idx[key]][:140]}\n")
            try:
                hm = float(r[idx[next(h for h in hdr if h.endswith('sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg'))]].replace(",", ""))
                act = float(r[idx["sm__cycles_active.avg"]].replace(",", ""))
                f.write(f"{'derived: tensor-pipe active = hmma_cycles_active / 4 sub-cores / sm__cycles_active':75s} {'%':14s} {hm / 4 / act * 100:.1f}\n")
            except Exception:
                pass
            f.write("\n")
    print(open(out).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
