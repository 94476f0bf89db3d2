"""prints the measured integer-pipe peaks (POPC + IADD3, and the LOP3 + POPC + IADD3 Hamming triple) of the GPU"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam3_comments_ghr_b200 import matcher  # noqa: E402

ctx = matcher.Context(0)
out = {}
for v, name in ((0, "popc_iadd3"), (1, "lop3_popc_iadd3")):
    r = [ctx.measure_popc_peak(v) for _ in range(3)]
    best = max(r, key=lambda x: x["popc_per_s"])
    out[name] = {"tpopc32_per_s": best["popc_per_s"] / 1e12, "popc32_per_clk_per_sm": best["per_clk_sm"], "sm_mhz": best["sm_mhz"],
                 "all": [x["popc_per_s"] / 1e12 for x in r]}
print(json.dumps(out, indent=1))
