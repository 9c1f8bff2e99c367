"""Whole config-2 batch (1 Mi tile pairs, seed 42) against the oracle: all six result fields and every traceback
state of EVERY tile (VERDICT r1, item 9 / SURVEY 8d "bit-exact on every tile vs L1").  The oracle runs on all host
cores in chunks; one JSON record is written (kept under profiles/).

  python tools/full_batch_parity.py [n_tiles] [out.json]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "darwin-gpu_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import oracle as O
import pygact as G
import synth
from helpers import compare_batch, engine_descs, oracle_descs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "full_batch_parity.json")
mb = synth.tile_microbatch(n, tile_size=320, seed=42)
with G.GactEngine(max_tiles=n) as eng:
    eng.upload(G.SET_REF, [mb["ref"].tobytes()])
    eng.upload(G.SET_READS, [mb["query"].tobytes()])
    eng.stage(engine_descs(G, mb))
    eng.run_staged()
    res, st = eng.fetch_staged()
    variant = eng.get_kernel()
od = oracle_descs(O, mb)
cores = len(os.sched_getaffinity(0))
bad_total, t0, CH = 0, time.time(), 1 << 16
first_bad = []
for lo in range(0, n, CH):
    hi = min(n, lo + CH)
    ores, ost = O.align_batch(mb["ref"], mb["query"], od[lo:hi], n_threads=cores)
    bad = compare_batch(res[lo:hi], st[lo:hi], ores, ost)
    bad_total += len(bad)
    if len(bad) and not first_bad:
        first_bad = [int(lo + b) for b in bad[:5]]
rec = {"what": "config 2, every tile: score, max_i, max_j, n_states, i_steps, j_steps and all traceback states, GPU (C ABI, "
               "stage/run_staged/fetch_staged) vs oracle/gact_oracle.c", "tiles": n, "kernel_variant": variant,
       "cells": int((mb["ref_len"].astype(np.int64) * mb["query_len"]).sum()), "states_compared": int(res["n_states"].sum()),
       "first_tiles": int(mb["first"].sum()), "mismatching_tiles": bad_total, "first_mismatches": first_bad,
       "oracle_seconds": round(time.time() - t0, 1), "oracle_threads": cores}
os.makedirs(os.path.dirname(out_path), exist_ok=True)
json.dump(rec, open(out_path, "w"), indent=1)
print(json.dumps(rec))
sys.exit(1 if bad_total else 0)
