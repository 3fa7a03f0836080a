"""Summarise an ncu launch list (gpu__time_duration.sum per launch, csv) by kernel name."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
    rows.append((name, us, r.get("Grid Size", ""), r.get("Block Size", "")))
tot = sum(u for _, u, _, _ in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, u, _, _ in rows:
    agg[n][0] += 1
    agg[n][1] += u
print(f"{len(rows)} launches, {tot/1e3:.3f} ms total (cold-cache, serialised under ncu)")
print(f"{'kernel':60s} {'n':>5s} {'ms':>9s} {'share':>7s} {'avg us':>9s}")
for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:60]:60s} {c:5d} {u/1e3:9.3f} {100*u/tot:6.1f}% {u/c:9.1f}")
if len(sys.argv) > 2 and sys.argv[2] == "--list":
    for i, (n, u, g, b) in enumerate(rows):
        print(i, n[:40], f"{u:.1f}us", g, b)
