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
print("# per kernel: launches, mean and median duration, share of the listed launches' time, mean DRAM bytes read / written per launch")
for k, m in sorted(agg.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", []))):
    t = m.get("gpu__time_duration.sum", [])
    rd, wr = m.get("dram__bytes_read.sum", []), m.get("dram__bytes_write.sum", [])
    extra = ""
    if rd and wr:
        extra = "  dram read %8.1f MB  write %8.1f MB" % (sum(rd) / len(rd) / 1e6, sum(wr) / len(wr) / 1e6)
    med = sorted(t)[len(t) // 2]
    print("%-58s launches %4d  mean %9.1f us  median %9.1f us  share %.3f%s" % (k, len(t), sum(t) / len(t) / 1e3, med / 1e3, sum(t) / tot, extra))
if any("dram__bytes_read.sum" in m for m in agg.values()):
    total = sum(sum(m.get("dram__bytes_read.sum", [])) + sum(m.get("dram__bytes_write.sum", [])) for m in agg.values())
    print("# total DRAM traffic of the listed launches: %.3f GB" % (total / 1e9))

# the last timed step of the bench command: the launches after the last k_aux_bound / k_aux_keys launch that is followed by >= 13 sweep launches
rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
names = [r["Kernel Name"] for r in rows]
starts = [i for i, n in enumerate(names) if n.startswith("void k_aux_bound")]
for a, b in zip(starts, starts[1:] + [len(names)]):
    seg = rows[a:b]
    if sum(1 for r in seg if "k_sweep_tc16" in r["Kernel Name"]) == 13:
        last = seg
step = collections.defaultdict(float)
for r in last:
    if any(x in r["Kernel Name"] for x in ("k_aux_bound", "k_sweep_tc16", "k_gather_z", "k_pre_", "k_scan_order")):
        step[r["Kernel Name"][:40]] += float(r["Metric Value"].replace(",", ""))
tt = sum(step.values())
print("# one stationary sweep (the last one with 13 sweep-kernel launches), serialised cold-cache durations under ncu:")
for k, v in sorted(step.items(), key=lambda kv: -kv[1]):
    print("#   %-42s %9.3f ms  share %.3f" % (k, v / 1e6, v / tt))
