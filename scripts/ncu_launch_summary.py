"""per-kernel totals of an `ncu --metrics ... --csv` launch list (one row per launch and metric)"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for row in csv.DictReader(lines):
    try:
        agg[row["Kernel Name"][:56]][row["Metric Name"]].append(float(row["Metric Value"].replace(",", "")))
    except (ValueError, KeyError):
        pass
tot = sum(sum(m.get("gpu__time_duration.sum", [])) for m in agg.values())
print("# per kernel: launches, mean duration, share of the listed launches' time, mean DRAM bytes read / written per launch")
for k, m in sorted(agg.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", []))):
    t = m.get("gpu__time_duration.sum", [])
    rd, wr = m.get("dram__bytes_read.sum", []), m.get("dram__bytes_write.sum", [])
    extra = ""
    if rd and wr:
        extra = "  dram read %8.1f MB  write %8.1f MB" % (sum(rd) / len(rd) / 1e6, sum(wr) / len(wr) / 1e6)
    print("%-58s launches %4d  mean %9.1f us  share %.3f%s" % (k, len(t), sum(t) / len(t) / 1e3, sum(t) / tot, extra))
if any("dram__bytes_read.sum" in m for m in agg.values()):
    total = sum(sum(m.get("dram__bytes_read.sum", [])) + sum(m.get("dram__bytes_write.sum", [])) for m in agg.values())
    print("# total DRAM traffic of the listed launches: %.3f GB" % (total / 1e9))
