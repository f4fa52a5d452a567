"""profiles/*_launches.csv (ncu --metrics gpu__time_duration.sum --csv) -> markdown table of kernels by total time."""
import csv
import re
import sys
from collections import defaultdict

src, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("spk::", "").replace("void ", "").strip()
    name = re.sub(r"\(bool\)|\(int\)", "", name)
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
print(title)
print()
print("| kernel | launches | total ms | share |")
print("|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v / total >= 0.001:
        print("| `%s` | %d | %.3f | %.1f %% |" % (k[:90], cnt[k], v, 100 * v / total))
print("| (all %d launches) | | %.3f | |" % (sum(cnt.values()), total))
