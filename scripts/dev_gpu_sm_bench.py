"""Developer check under gpurun: bench.py's split-merge figures alone."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import noparama_b200 as npb
from noparama_b200 import synthetic as syn
ctx = npb.Context(0)
X, y = syn.config(3)
ds = npb.Dataset(ctx, X)
npb.NormalInverseWishart(**syn.reference_prior(16)).bind(ctx)
t0 = time.time()
r = bench.split_merge_measure(npb, syn, ctx, ds, X, y, 74.0, 0, cpu=False)
print("seconds", time.time() - t0)
for reg in (r, r["accepting_regime_2d"]):
    for k in ("jain_neal", "triadic"):
        print(k, {a: b for a, b in reg[k].items() if a != "roofline"}, "frac %.3f" % reg[k]["roofline"]["frac"])
