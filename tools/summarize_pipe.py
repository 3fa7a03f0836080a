"""Whole-step tensor-pipe table from an ncu metrics list (csv log of
  --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum):
one row per launch (out csv) and the time-weighted tensor-pipe activity over the step, per kernel and in total."""
import csv
import re
import sys
from collections import defaultdict

src, out_csv = sys.argv[1], sys.argv[2]
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
per = defaultdict(dict)
order = []
for r in csv.DictReader(lines):
    i = int(r["ID"])
    if i not in per:
        order.append(i)
        per[i]["kernel"] = re.sub(r"\(.*", "", r["Kernel Name"]).strip().replace("void ", "").replace("nope::", "")
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    n = r["Metric Name"]
    if n == "gpu__time_duration.sum":
        per[i]["us"] = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    elif n.startswith("sm__pipe_tensor_cycles_active"):
        per[i]["tensor_pct"] = v
    elif n == "dram__bytes_read.sum":
        per[i]["rd_MB"] = v / 1e6 if u == "byte" else (v if u == "Mbyte" else v * (1e3 if u == "Gbyte" else 1e-3))
    elif n == "dram__bytes_write.sum":
        per[i]["wr_MB"] = v / 1e6 if u == "byte" else (v if u == "Mbyte" else v * (1e3 if u == "Gbyte" else 1e-3))
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch", "kernel", "dur_us", "tensor_pipe_active_pct", "dram_read_MB", "dram_write_MB"])
    for k, i in enumerate(order):
        d = per[i]
        w.writerow([k, d["kernel"][:44], f"{d.get('us', 0):.1f}", f"{d.get('tensor_pct', 0):.1f}",
                    f"{d.get('rd_MB', 0):.1f}", f"{d.get('wr_MB', 0):.1f}"])
tot = sum(per[i].get("us", 0) for i in order)
tw = sum(per[i].get("us", 0) * per[i].get("tensor_pct", 0) for i in order) / tot
agg = defaultdict(lambda: [0, 0.0, 0.0])
for i in order:
    a = agg[per[i]["kernel"]]
    a[0] += 1
    a[1] += per[i].get("us", 0)
    a[2] += per[i].get("us", 0) * per[i].get("tensor_pct", 0)
print(f"{len(order)} launches, {tot / 1e3:.3f} ms (serialised under ncu); time-weighted tensor pipe active over the step: {tw:.1f} %")
print(f"{'kernel':48s} {'n':>4s} {'ms':>8s} {'share':>7s} {'tensor %':>9s}")
for n, (c, u, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:48]:48s} {c:4d} {u / 1e3:8.3f} {100 * u / tot:6.1f}% {t / u if u else 0:9.1f}")
sweep = [i for i in order if not re.search(r"enc_|<128|<64|Cat|iota", per[i]["kernel"])]
ts = sum(per[i].get("us", 0) for i in sweep)
print(f"UNet sweep only (without the template encoder's launches): {ts / 1e3:.3f} ms, tensor pipe "
      f"{sum(per[i].get('us', 0) * per[i].get('tensor_pct', 0) for i in sweep) / ts:.1f} %")
