"""Developer check under gpurun: FP32 peak probes (scalar FFMA vs packed FFMA2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NPB_PEAK_VERBOSE"] = "1"
import noparama_b200 as npb
from bench import mc_fp32_peak
ctx = npb.Context(0)
print("npb_fp32_peak ->", mc_fp32_peak(ctx))
