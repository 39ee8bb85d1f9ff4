"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, rows = r, rows[i + 1:]
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows:
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
    a = agg.setdefault(r[ki][:70], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{a[1]:10.1f} us {100*a[1]/tot:5.1f}%  {a[0]:5d} launches  {a[1]/a[0]:8.1f} us/launch  {n}")
print(f"total {tot:.1f} us")
