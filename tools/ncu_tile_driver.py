"""Three launches of the packed tile kernel on a 131 072-tile config-2 batch (for ncu captures):
    python tools/ncu_tile_driver.py [n_tiles]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "darwin-gpu_b200"))
import numpy as np
import pygact as G
import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
mb = synth.tile_microbatch(n, tile_size=320, seed=42)
with G.GactEngine(max_tiles=n) as eng:
    eng.upload(G.SET_REF, [mb["ref"].tobytes()])
    eng.upload(G.SET_READS, [mb["query"].tobytes()])
    d = G.make_descs(n)
    for k in ("ref_off", "query_off", "ref_len", "query_len", "reverse", "first"):
        d[k] = mb[k]
    d["ref_set"], d["query_set"] = G.SET_REF, G.SET_READS
    eng.stage(d)
    for _ in range(3):
        eng.run_staged()
    eng.sync()
    print("kernel_ms", eng.last_kernel_ms(), "tiles", n)
    eng.fetch_staged()
