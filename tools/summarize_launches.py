"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
per kernel name -- launches, total / share / average duration, DRAM MB.  usage: summarize_launches.py in.csv [title]"""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
i_name, i_metric, i_val, i_id = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
i_unit = hdr.index("Metric Unit")
per = collections.OrderedDict()
for r in rows[1:]:
    k = (r[i_id], r[i_name])
    d = per.setdefault(k, {})
    v = float(r[i_val].replace(",", ""))
    u = r[i_unit]
    if r[i_metric].startswith("gpu__time"):
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        d["us"] = v
    else:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        d["bytes"] = d.get("bytes", 0.0) + v
agg = {}
for (_, name), d in per.items():
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("us", 0.0); a[2] += d.get("bytes", 0.0)
tot = sum(a[1] for a in agg.values())
print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# (per-launch times under ncu are cold-cache and serialised: use the SHARES)")
print("# %d launches, %.2f ms" % (sum(a[0] for a in agg.values()), tot / 1e3))
print("#\n# %-70s %6s %10s %6s %9s %10s" % ("kernel", "calls", "total us", "share", "avg us", "dram MB"))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-70s %6d %10.1f %5.1f%% %9.1f %10.1f" % (name[:70], a[0], a[1], 100 * a[1] / tot, a[1] / a[0], a[2] / 1e6))
