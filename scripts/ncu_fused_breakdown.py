"""Instruction-level breakdown of an `ncu --set full --import-source on` capture of k_sweep_tc16: executed warp instructions per
(chain, step) by opcode and by contiguous SASS segment (segments = runs of instructions with similar execution counts, i.e. the
warp roles' loops), with their share of the stall samples.  usage: python scripts/ncu_fused_breakdown.py report.ncu-rep chains steps"""
import collections
import csv
import subprocess
import sys

rep, chains, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return 0.0


units = float(chains) * steps
tot_i = sum(f(r, "Instructions Executed") for r in data)
tot_s = sum(f(r, "# Samples") for r in data)
print("kernel: %s" % rows[0][1])
print("chain-steps in the launch: %d x %d; executed warp instructions per chain-step: %.2f" % (chains, steps, tot_i / units))
ops, smp = collections.Counter(), collections.Counter()
for r in data:
    t = r[ix["Source"]].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += f(r, "Instructions Executed")
    smp[op] += f(r, "# Samples")
print("\nby opcode (warp instructions per chain-step, share of stall samples):")
for op, v in ops.most_common(24):
    print("  %-10s %6.2f  %5.1f%%" % (op, v / units, 100 * smp[op] / tot_s))
print("\nby SASS segment (index range, instructions, executions per instruction, warp instructions per chain-step, samples):")
seg, cur = [], None
for n, r in enumerate(data):
    e = f(r, "Instructions Executed")
    if cur is None or not (0.5 * cur["e"] <= e <= 2 * cur["e"]):
        cur = {"a": n, "e": max(e, 1.0), "i": 0.0, "s": 0.0, "n": 0, "ops": collections.Counter()}
        seg.append(cur)
    cur["i"] += e
    cur["s"] += f(r, "# Samples")
    cur["n"] += 1
    cur["b"] = n
    t = r[ix["Source"]].split()
    cur["ops"][(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += 1
for s in seg:
    if s["i"] / tot_i > 0.004 or s["s"] / tot_s > 0.004:
        top = " ".join("%s:%d" % kv for kv in s["ops"].most_common(5))
        print("  %5d-%5d n=%4d x %.2e  %6.2f  %5.1f%%  %s" % (s["a"], s["b"], s["n"], s["i"] / s["n"], s["i"] / units, 100 * s["s"] / tot_s, top))
