"""Join an ncu --page raw csv of the LDM sweep's conv_tc2 launches with the layer schedule
(same order as nope_ldm::forward_chunk) -> per-launch TFLOP/s table."""
import csv
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from nope_b200.synth_weights import ldm_block_plan

N = int(sys.argv[2]) if len(sys.argv) > 2 else 128


def schedule(n_hyp, n_ref=1):
    inp, mid, out = ldm_block_plan()
    L = []

    def res(p, cin, cout, S, n):
        L.append((p + ".c1", n * S * S, cout, 9 * cin))
        L.append((p + ".c2", n * S * S, cout, 9 * cout + (cin if cin != cout else 0)))

    def st_pre(p, c, S, n):
        L.append((p + ".proj_in", n * S * S, c, c))
        L.append((p + ".qkv", n * S * S, 3 * c, c))
        L.append((p + ".to_out", n * S * S, c, c))

    def st_post(p, c, S, n):
        L.append((p + ".ff1g", n * S * S, 8 * c, c))
        L.append((p + ".ff2", n * S * S, c, 4 * c))
        L.append((p + ".proj_out", n * S * S, c, c))

    res("input_blocks.1.0", 256, 256, 32, n_ref)
    st_pre("input_blocks.1.1", 256, 32, n_ref)
    S, n = 32, n_hyp
    for i, b in enumerate(inp):
        p = f"input_blocks.{i}"
        if i == 1:
            st_post(p + ".1", 256, S, n)
        elif b[0] == "res":
            res(p + ".0", b[1], b[2], S, n)
            st_pre(p + ".1", b[2], S, n)
            st_post(p + ".1", b[2], S, n)
        elif b[0] == "down":
            S //= 2
            L.append((p + ".0.op", n * S * S, b[2], 9 * b[1]))
    res("middle_block.0", mid, mid, S, n)
    st_pre("middle_block.1", mid, S, n); st_post("middle_block.1", mid, S, n)
    res("middle_block.2", mid, mid, S, n)
    for i, b in enumerate(out):
        p = f"output_blocks.{i}"
        res(p + ".0", b[1], b[2], S, n)
        st_pre(p + ".1", b[2], S, n); st_post(p + ".1", b[2], S, n)
        if b[4]:
            L.append((p + ".2.conv", n * S * S, 4 * b[2], 4 * b[2]))   # folded: 4 parity GEMMs, K = 4 Cin
            S *= 2
    L.append(("out.2", n * S * S, 64, 9 * 256))
    return L


rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
sched = schedule(N)
print(f"{'layer':32s} {'M':>8s} {'N':>5s} {'K':>6s} {'us':>8s} {'TFLOP/s':>8s} {'tensor%':>8s} {'dram GB':>8s}")
tot_f = tot_t = 0
for r, (name, M, Nn, K) in zip(rows[2:], sched):
    us = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
    unit = rows[1][idx["gpu__time_duration.sum"]]
    if unit.startswith("ns"): us /= 1e3
    if unit.startswith("ms"): us *= 1e3
    fl = 2.0 * M * Nn * K
    tp = r[idx["sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"]] if "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active" in idx else ""
    def gb(k):
        v = float(r[idx[k]].replace(",", "")); u = rows[1][idx[k]]
        return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1e-9)
    d = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
    print(f"{name:32s} {M:8d} {Nn:5d} {K:6d} {us:8.1f} {fl / us / 1e6:8.0f} {tp:>8s} {d:8.3f}")
    tot_f += fl; tot_t += us
print(f"total {tot_t / 1e3:.2f} ms, {tot_f / tot_t / 1e6:.0f} TFLOP/s over {min(len(rows) - 2, len(sched))} launches")
