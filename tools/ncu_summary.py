"""Per-kernel table out of an ncu report (raw page CSV on stdin or a path): launches, mean duration, DRAM bytes per launch,
achieved DRAM GB/s and the share of the measured HBM peak (MEASURED_PEAKS.json when present).
  ncu -i gpurun_out/x.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/r02_small_kernels.md"""
import collections, csv, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6574.5
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)
except Exception:
    pass
rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def col(r, name, unit_scale=None):
    i = ix.get(name)
    if i is None or r[i] in ("", "n/a"):
        return None
    v = float(r[i].replace(",", ""))
    u = units[i]
    scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}.get(u, 1.0)
    return v * scale
agg = collections.OrderedDict()
for r in data:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("<unnamed>::", "")
    t = col(r, "gpu__time_duration.sum")
    rd, wr = col(r, "dram__bytes_read.sum") or 0.0, col(r, "dram__bytes_write.sum") or 0.0
    a = agg.setdefault(name, {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0, "sm": 0.0, "regs": None, "l2": 0.0})
    a["n"] += 1; a["t"] += t; a["rd"] += rd; a["wr"] += wr
    a["sm"] += col(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed") or 0.0
    a["l2"] += col(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed") or 0.0
    a["regs"] = r[ix["launch__registers_per_thread"]] if "launch__registers_per_thread" in ix else None
print(f"| kernel | launches | mean us | DRAM read MB / launch | DRAM write MB / launch | DRAM GB/s | share of {peak:.0f} GB/s | SM % | L2 % | regs |")
print("|---|---|---|---|---|---|---|---|---|---|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    n = a["n"]
    gbs = (a["rd"] + a["wr"]) / a["t"] / 1e9 if a["t"] > 0 else 0.0
    print(f"| {name} | {n} | {a['t'] / n * 1e6:.1f} | {a['rd'] / n / 1e6:.2f} | {a['wr'] / n / 1e6:.2f} | {gbs:.0f} | {gbs / peak:.2f} | "
          f"{a['sm'] / n:.0f} | {a['l2'] / n:.0f} | {a['regs']} |")
