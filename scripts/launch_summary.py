"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: total ms, launches, mean us, share."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    val = float(d["Metric Value"].replace(",", ""))
    val = val / 1e3 if d["Metric Unit"] == "us" else val / 1e6 if d["Metric Unit"] == "ns" else val
    k = re.sub(r"\(.*", "", d["Kernel Name"])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += val
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.3f} ms {v[0]:5d} x {v[1] / v[0] * 1e3:8.1f} us  {100 * v[1] / tot:5.1f}%  {k[:100]}")
print(f"{tot:9.3f} ms total")
