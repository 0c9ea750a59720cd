"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import collections, csv, io, sys
lines = open(sys.argv[1]).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
    if row["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(row["Metric Unit"], 1.0)
    k = row["Kernel Name"][:90]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{sum(v[0] for v in agg.values())} launches, {tot:.1f} us of kernel time")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:5d}  avg={v[1] / v[0]:8.2f} us  {k}")
