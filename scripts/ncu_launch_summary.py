"""per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    try:
        agg[row["Kernel Name"][:56]].append(float(row["Metric Value"].replace(",", "")))
    except (ValueError, KeyError):
        pass
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-58s launches %4d  mean %9.1f us  share %.3f" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / tot))
