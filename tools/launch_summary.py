"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (share of the step)."""
import collections
import csv
import io
import sys

txt = open(sys.argv[1]).read()
rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    n = r["Kernel Name"].split("(")[0].replace("void ", "").replace("cnb::", "")
    n = n if len(n) < 70 else n[:67] + "..."
    agg[n][0] += 1
    agg[n][1] += float(r["Metric Value"]) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot:.2f} ms summed kernel time")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s}{v[0]:5d}{v[1]:10.3f} ms {100 * v[1] / tot:6.1f}%")
if len(sys.argv) > 2:
    print()
    for r in sorted(rows, key=lambda r: -float(r["Metric Value"]))[: int(sys.argv[2])]:
        print(r["ID"], r["Kernel Name"][:60], r["Grid Size"], float(r["Metric Value"]) / 1e6)
