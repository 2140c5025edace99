"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per (kernel, grid): python tools/launch_summary.py file.csv"""
import collections, csv, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].replace("unnamed>::", "").replace("void ", "")
    name = name.split("(")[0][-44:]
    a = agg.setdefault((name, r[8]), [0, 0.0])
    a[0] += 1
    a[1] += float(r[14]) / 1e3
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':46s} {'grid':16s} {'n':>4s} {'total us':>10s} {'avg us':>9s} {'share':>6s}")
for (k, g), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:46s} {g:16s} {n:4d} {t:10.1f} {t / n:9.1f} {t / tot:6.1%}")
print(f"total {tot:.1f} us over {len(rows)} launches")
