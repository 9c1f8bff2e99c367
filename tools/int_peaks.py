"""Prints the integer / DPX issue-rate microbenchmark (gact_int_peak) for every instruction kind."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "darwin-gpu_b200"))
import pygact as G  # noqa: E402

KINDS = ["IADD3", "VIMNMX3.S32", "VIADDMNMX.S32", "VIMNMX3.S16x2", "VIADDMNMX.S16x2", "LOP3", "IMAD",
         "VIADDMNMX+IMAD 1:1", "HSET2", "VIADD.16x2", "PRMT", "VIMNMX3.S16x2+HSET2 1:1", "SHFL"]
out = {}
for k, name in enumerate(KINDS):
    out[name] = round(G.int_peak(k), 1)
    print(f"{name:28s} {out[name]:10.1f} G lane-ops/s")
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
