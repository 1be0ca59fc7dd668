"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(float); cnt = defaultdict(int)
per_launch = []
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    tot[name] += us; cnt[name] += 1
    per_launch.append((us, name, r.get("Grid Size", ""), r.get("ID", "")))
total = sum(tot.values())
print(f"# total {total:.1f} us over {sum(cnt.values())} launches")
for name, us in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{us:10.1f} us  {100 * us / total:5.1f}%  n={cnt[name]:4d}  {name[:90]}")
print("# 25 longest launches")
for us, name, grid, i in sorted(per_launch, reverse=True)[:25]:
    print(f"{us:10.1f} us  id={i} grid={grid} {name[:80]}")
