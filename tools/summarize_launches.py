import csv, sys, collections, re
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = f.readlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rd = csv.DictReader(lines[start:])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum": continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if unit in ("nsecond", "ns"): v /= 1e3
    elif unit in ("msecond", "ms"): v *= 1e3
    elif unit in ("second", "s"): v *= 1e6
    rows.append((r["Kernel Name"], v))
tot = sum(v for _, v in rows)
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in rows:
    k = re.sub(r"<.*", "", k)[:70]
    agg[k][0] += 1; agg[k][1] += v
print("launches %d  total %.1f us" % (len(rows), tot))
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%7.1f us %5.1f%%  x%-4d avg %6.1f  %s" % (v, 100 * v / tot, n, v / n, k))
