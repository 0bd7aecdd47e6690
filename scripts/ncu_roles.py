"""Splits the stall samples of an ncu report of k_alg8_sweep_tile4 into producer / consumer (before / after the
USETMAXREG.DEALLOC) and lists the hottest instructions.  usage: python scripts/ncu_roles.py report.ncu-rep [chain_steps]"""
import csv, subprocess, sys
rep = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 2.368e8
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]].replace(",", ""))
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
cons = next(i for i, r in enumerate(data) if "USETMAXREG.DEALLOC" in r[ix["Source"]])
keys = ["stall_barrier", "stall_wait", "stall_long_sb", "stall_short_sb", "stall_math", "stall_not_selected", "stall_selected",
        "stall_dispatch", "stall_mio", "stall_branch_resolving", "stall_no_inst", "stall_lg"]
for name, (a, b) in {"producer": (0, cons), "consumer": (cons, len(data))}.items():
    seg = data[a:b]
    ins = sum(f(r, "Instructions Executed") for r in seg)
    print("%s: samples %.1f%%, %.1f instructions per chain-step" % (name, 100 * sum(f(r, "# Samples") for r in seg) / tot, ins / steps))
    print("   " + " ".join("%s %.1f%%" % (k[6:], 100 * sum(f(r, k) for r in seg) / tot) for k in keys))
    for r in sorted(seg, key=lambda r: -f(r, "# Samples"))[:12]:
        print("   %6.2f%% exec %.2e  %s" % (100 * f(r, "# Samples") / tot, f(r, "Instructions Executed"), r[ix["Source"]][:80]))
