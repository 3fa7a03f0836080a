"""Compact per-launch summary of an `ncu --page raw --csv` export (conv or memory kernels)."""
import csv
import json
import sys

src, out_csv = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("ID", "id"), ("Kernel Name", "kernel"), ("launch__grid_size", "grid"),
        ("gpu__time_duration.sum", "dur_us"),
        ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_mem_active_pct"),
        ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
        ("lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("launch__registers_per_thread", "regs"),
        ("sm__cycles_elapsed.avg.per_second", "sm_ghz")]
cols = [(c, n) for c, n in cols if c in ix]
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([n for _, n in cols])
    for r in data:
        if len(r) < len(hdr) - 2:
            continue
        w.writerow([(r[ix[c]].split("(")[0][:48] if n == "kernel" else r[ix[c]]) for c, n in cols])
tot_t = sum(float(r[ix["gpu__time_duration.sum"]]) for r in data if len(r) >= len(hdr) - 2)
tot_b = sum(float(r[ix["dram__bytes_read.sum"]]) + float(r[ix["dram__bytes_write.sum"]])
            for r in data if len(r) >= len(hdr) - 2)
n = sum(1 for r in data if len(r) >= len(hdr) - 2)
print(json.dumps({"launches": n, "total_us": tot_t, "dram_MB_total": tot_b,
                  "dram_bytes_per_launch": tot_b * 1e6 / n if n else None}))
