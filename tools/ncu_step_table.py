"""`ncu -i <rep> --page raw --csv` of one training step -> markdown table (one row per launch, in launch order).
Usage: python tools/ncu_step_table.py gpurun_out/r02_step_raw.csv [title]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
COLS = [("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "rd MB"), ("dram__bytes_write.sum", "wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 %"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs")]
cols = [(k, t) for k, t in COLS if k in ix]


def num(r, k):
    v = r[ix[k]].replace(",", "")
    try:
        x = float(v)
    except ValueError:
        return v
    u = units[ix[k]]
    if u == "Gbyte":
        x *= 1e3
    elif u == "Kbyte":
        x /= 1e3
    elif u == "byte":
        x /= 1e6
    elif u in ("ms", "msecond"):
        x *= 1e3
    elif u in ("ns", "nsecond"):
        x /= 1e3
    return "%.4g" % x


TAGS = sys.argv[3].split(",") if len(sys.argv) > 3 else []
STALLS = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]


def top_stalls(r):
    vals = []
    for h in STALLS:
        try:
            vals.append((float(r[ix[h]].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        except ValueError:
            pass
    vals.sort(reverse=True)
    return ", ".join("%s %.1f" % (n, v) for v, n in vals[:3])


def dram_mb(r):
    k = "dram__bytes.sum.per_second"
    if k not in ix:
        return ""
    v = float(r[ix[k]].replace(",", ""))
    u = units[ix[k]]
    v *= {"Tbyte/s": 1e12, "Gbyte/s": 1e9, "Mbyte/s": 1e6, "Kbyte/s": 1e3}.get(u, 1.0)
    t = float(r[ix["gpu__time_duration.sum"]].replace(",", ""))
    t *= {"us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "ns": 1e-9, "nsecond": 1e-9}[units[ix["gpu__time_duration.sum"]]]
    return "%.1f" % (v * t / 1e6)


print(sys.argv[2] if len(sys.argv) > 2 else "")
print()
print("| # | kernel | " + " | ".join(t for _, t in cols) + " | dram MB | top warp stalls (warps per issue) |")
print("|---|---|" + "---|" * (len(cols) + 2))
tot = 0.0
for i, r in enumerate(data):
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("spk::", "").replace("void ", "")
    tag = (" " + TAGS[i]) if i < len(TAGS) else ""
    print("| %d | `%s`%s | " % (i, name[:48], tag) + " | ".join(num(r, k) for k, _ in cols) + " | %s | %s |" % (dram_mb(r), top_stalls(r)))
    tot += float(num(r, "gpu__time_duration.sum"))
print()
print("sum of gpu__time_duration: %.1f us over %d launches (cold caches, serialised)" % (tot, len(data)))
