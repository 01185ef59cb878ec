"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
    python scripts/summarize_launches.py launches.csv [skip_launches]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = None
tot = collections.Counter()
cnt = collections.Counter()
n = 0
for r in rows:
    if hdr is None:
        if "Kernel Name" in r and "Metric Value" in r:
            hdr = r
        continue
    if len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    n += 1
    if n <= skip:
        continue
    name = re.sub(r"\(.*", "", d["Kernel Name"])
    name = re.sub(r"^void (stfem::)?", "", name)
    v = float(d["Metric Value"].replace(",", ""))
    unit = d.get("Metric Unit", "ns")
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print("%d launches, %.1f us in kernels" % (sum(cnt.values()), total))
for k, v in tot.most_common(25):
    print("%7.1f us %5.1f%% %6d x %8.2f us  %s" % (v, 100 * v / total, cnt[k], v / cnt[k], k[:110]))
