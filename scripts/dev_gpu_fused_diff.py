import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import noparama_b200 as npb
from noparama_b200 import synthetic as syn
from test_gpu_fused16 import overlapping, run
D = 16
ctx = npb.Context(0)
for chains, N, block, sweeps in ((7, 1013, 256, 1), (7, 1013, 256, 3), (4, 256, 128, 1), (2, 64, 128, 1), (2, 64, 128, 2)):
    os.environ["NPB_D16_BLOCK"] = str(block)
    X, y, means = overlapping(N, 6, 5)
    ds = npb.Dataset(ctx, X)
    out = {}
    for name, path, spec in (("fused", None, "1"), ("fused_seq", None, "0"), ("pair", "tc2", "1"), ("pair_seq", "tc2", "0")):
        if path:
            os.environ["NPB_D16_PATH"] = path
        else:
            os.environ.pop("NPB_D16_PATH", None)
        os.environ["NPB_D64_SPEC"] = spec
        mc, z, tot = run(npb, ctx, ds, means, chains, sweeps, seed=11)
        out[name] = (z, tot)
    print("chains", chains, "N", N, "block", block, "sweeps", sweeps)
    for a in out:
        for b in out:
            if a < b:
                d = np.argwhere(out[a][0] != out[b][0])
                print("  %s vs %s: %d differ; moved %d / %d; first %s" % (a, b, len(d), out[a][1]["moved"], out[b][1]["moved"], d[:4].tolist()))
    ds.close()
