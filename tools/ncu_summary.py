"""Per-kernel summary of an `ncu --csv --metrics ...` log: one row per kernel name (first instance) with every metric.
usage: python tools/ncu_summary.py file.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    key = (r[ki].split("(")[0][-60:], r[ii])
    per.setdefault(key, {})[r[mi]] = float(r[vi].replace(",", ""))
seen = set()
short = {"gpu__time_duration.sum": "us", "dram__bytes_read.sum": "rd_MB", "dram__bytes_write.sum": "wr_MB",
         "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor%",
         "sm__issue_active.avg.pct_of_peak_sustained_active": "issue%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1%",
         "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2%", "sm__warps_active.avg.pct_of_peak_sustained_active": "occ%",
         "launch__registers_per_thread": "regs", "smsp__inst_executed.sum": "Minst"}
print(f"{'kernel':62s} " + " ".join(f"{v:>9s}" for v in short.values()) + "   GB/s(rd+wr)")
for (name, _), m in per.items():
    if name in seen:
        continue
    seen.add(name)
    vals = []
    for k, s in short.items():
        v = m.get(k, float("nan"))
        if s == "us": v /= 1e3
        if s in ("rd_MB", "wr_MB"): v /= 1e6
        if s == "Minst": v /= 1e6
        vals.append(v)
    t_us, rd, wr = vals[0], vals[1], vals[2]
    print(f"{name:62s} " + " ".join(f"{v:9.1f}" for v in vals) + f"   {(rd + wr) / t_us * 1e3:9.0f}")
