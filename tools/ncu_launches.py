"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X)
into a per-kernel table: launches, total / average duration, share of the total.
Usage: python tools/ncu_launches.py launches.csv "title" [out.md]"""
import csv, re, sys
from collections import OrderedDict

path, title = sys.argv[1], sys.argv[2]
lines = [l for l in open(path, errors="replace") if not l.startswith("==")]
rows = list(csv.reader(lines))
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
agg = OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ci["Kernel Name"]]
    m = re.search(r"(k_[a-z_0-9]+)(<[^>]*>)?", name)
    key = (m.group(1) + (m.group(2) or "")) if m else name[:40]
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    us = v / 1000.0 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1000.0
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += us
total = sum(a[1] for a in agg.values())
out = ["# %s" % title, "",
       "%d launches captured (cold-cache, serialised: compare shares, not absolutes); total %.1f us" % (
           sum(a[0] for a in agg.values()), total), "",
       "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("| %s | %d | %.1f | %.1f | %.3f |" % (k, n, us, us / n, us / total))
text = "\n".join(out) + "\n"
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(text)
print(text)
