"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list by kernel
name: launches, summed time, share of the listed time and (when present) DRAM bytes and the implied GB/s."""
import collections
import csv
import io
import sys

txt = open(sys.argv[1]).read()
rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
UNIT = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per_id = collections.OrderedDict()
for r in rows:
    e = per_id.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r.get("Grid Size", ""), "ms": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    m = r["Metric Name"]
    if m.startswith("gpu__time_duration"):
        e["ms"] = v
    elif m.startswith("dram__bytes_read"):
        e["rd"] = v
    elif m.startswith("dram__bytes_write"):
        e["wr"] = v
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for e in per_id.values():
    n = e["name"].split("(")[0].replace("void ", "").replace("cnb::", "")
    n = n if len(n) < 70 else n[:67] + "..."
    a = agg[n]
    a[0] += 1
    a[1] += e["ms"]
    a[2] += e["rd"]
    a[3] += e["wr"]
tot = sum(v[1] for v in agg.values())
print(f"{len(per_id)} launches, {tot:.2f} ms summed kernel time (ncu: serialised, cold cache)")
print(f"{'kernel':70s}{'n':>5s}{'ms':>10s}{'share':>8s}{'GB rd':>9s}{'GB wr':>9s}{'GB/s':>8s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = (v[2] + v[3]) / 1e9 / (v[1] / 1e3) if v[1] > 0 else 0.0
    print(f"{k:70s}{v[0]:5d}{v[1]:10.3f}{100 * v[1] / tot:7.1f}%{v[2] / 1e9:9.2f}{v[3] / 1e9:9.2f}{gbs:8.0f}")
if len(sys.argv) > 2:
    print()
    for i, e in sorted(per_id.items(), key=lambda kv: -kv[1]["ms"])[: int(sys.argv[2])]:
        print(i, e["name"][:60], e["grid"], e["ms"])
