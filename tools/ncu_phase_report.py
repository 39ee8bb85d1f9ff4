"""Aggregate `ncu --page source --print-source cuda,sass --csv` output by kernel phase.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv ; python tools/ncu_phase_report.py src.csv
"""
import bisect, csv, sys, os
path = sys.argv[1]
rows = list(csv.reader(open(path)))
fname = [a for a in sys.argv[2:] if a.endswith(".cuh")]
fname = fname[0] if fname else "match_update_kernels.cuh"
src_file = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fastace_b200/csrc", fname)
src = open(src_file).read().split("\n")
marks = [(i + 1, l.strip()[:70]) for i, l in enumerate(src) if l.strip().startswith("// ----")]
bounds = [m[0] for m in marks]
cur_file = None
agg, other, lines = {}, {}, {}
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1]; continue
    if len(r) > 7 and r[0] == "Line No":
        hdr = r; ie = r.index("Instructions Executed"); smp = r.index("# Samples"); continue
    if hdr is None or len(r) <= ie or r[2] != "-":
        continue
    try:
        ln, v, s = int(r[0]), int(r[ie]), int(r[smp])
    except ValueError:
        continue
    if cur_file and cur_file.endswith(fname):
        k = bisect.bisect_right(bounds, ln) - 1
        name = marks[k][1] if k >= 0 else "helpers (top of file)"
        if k < 0:
            name = "helper: " + src[ln - 1].strip()[:60]
        a = agg.setdefault(name, [0, 0]); a[0] += v; a[1] += s
        lines[ln] = lines.get(ln, 0) + v
    else:
        a = agg.setdefault("other file: " + os.path.basename(cur_file or "?"), [0, 0]); a[0] += v; a[1] += s
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print("total warp instructions", tot, " samples", tots)
for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%11d %5.1f%% inst  %5.1f%% samples  %s" % (v[0], 100 * v[0] / tot, 100 * v[1] / max(tots, 1), k))
if "--lines" in sys.argv:
    for ln, v in sorted(lines.items(), key=lambda x: -x[1])[:40]:
        print("%10d %5.1f%%  %4d: %s" % (v, 100 * v / tot, ln, src[ln - 1].strip()[:100]))
