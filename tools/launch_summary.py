"""Sum an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name: count, total us, share."""
import csv
import re
import sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = {}
for r in rows:
    if r is hdr or r[0] == "ID" or not r[0].isdigit():
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[ui], 1.0)
    name = re.sub(r"\(.*$", "", r[ki])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{t:10.1f} us {100 * t / tot:5.1f} %  x{n:<4d} {name[:150]}")
print(f"{tot:10.1f} us total")
