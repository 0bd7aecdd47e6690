"""profiles/ summary of the D = 64 path: launch list (gpu__time_duration per launch) + selected counters of one
`ncu --set full` capture.  usage: ncu_d64_summary.py launches.csv prof.ncu-rep > profiles/...txt"""
import collections, csv, subprocess, sys
launches, rep = sys.argv[1], sys.argv[2]
lines = [l for l in open(launches) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    try:
        agg[row["Kernel Name"][:48]].append(float(row["Metric Value"].replace(",", "")))
    except (ValueError, KeyError):
        pass
tot = sum(sum(v) for v in agg.values())
print("# launch list (ncu --metrics gpu__time_duration.sum --clock-control none): per-launch times are cold-cache and serialised")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-50s launches %4d  mean %9.1f us  share %.3f" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / tot))
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
        "sm__pipe_tensor_cycles_active.avg.pct", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__throughput.avg.pct", "gpu__dram_throughput.avg.pct", "sm__cycles_active.avg", "smsp__average_warps_issue_stalled",
        "sm__warps_active.avg.pct"]
kn = hdr.index("Kernel Name")
for vals in rows[2:]:
    print("\n# ncu --set full --clock-control none:", vals[kn][:70])
    for i, h in enumerate(hdr):
        if not any(w in h for w in want) or ".max" in h or ".min" in h or ("TriageCompute" in h):
            continue
        try:
            if "stalled" in h and float(vals[i].replace(",", "")) < 0.3:
                continue
        except ValueError:
            pass
        print("  %-95s %-14s %s" % (h, units[i], vals[i]))
