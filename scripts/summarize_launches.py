"""Summarise an ncu launch-list CSV (gpu__time_duration.sum plus optional metrics) per kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
per = collections.OrderedDict()
for r in rows:
    key = (r["ID"], re.sub(r"\(.*", "", r["Kernel Name"]))
    d = per.setdefault(key, {})
    v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
    unit = r["Metric Unit"]
    name = r["Metric Name"]
    if name == "gpu__time_duration.sum":
        v = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    if name.startswith("dram__bytes"):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        v = v * mult
    d[name] = v
agg = collections.OrderedDict()
tot = 0.0
for (_, k), d in per.items():
    a = agg.setdefault(k, {"n": 0, "ms": 0.0, "tens": 0.0, "bytes": 0.0})
    ms = d.get("gpu__time_duration.sum", 0.0)
    a["n"] += 1
    a["ms"] += ms
    a["tens"] += ms * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    a["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot += ms
print(f"total {tot:.2f} ms over {len(per)} launches")
print(f"{'ms':>9} {'share':>6} {'n':>4} {'avg us':>9} {'tensor%':>8} {'DRAM GB/s':>10}  kernel")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    gbs = a["bytes"] / (a["ms"] * 1e-3) / 1e9 if a["ms"] > 0 else 0
    print(f"{a['ms']:9.3f} {100*a['ms']/tot:5.1f}% {a['n']:4d} {1000*a['ms']/a['n']:9.1f} {a['tens']/max(a['ms'],1e-9):8.1f} {gbs:10.0f}  {k[:100]}")
