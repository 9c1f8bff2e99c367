"""Device-resident GCUPS of the tile kernels over the config-5 tile sizes (256/96, 320/120, 512/192, 1024/384):
    python tools/tile_sweep.py [cells_per_size, default 2e10]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "darwin-gpu_b200"))
import numpy as np
import pygact as G
import synth

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 2e10
for T, O in ((256, 96), (320, 120), (512, 192), (1024, 384)):
    n = int(budget / (T * T))
    mb = synth.tile_microbatch(n, tile_size=T, seed=7)
    cells = int((mb["ref_len"].astype(np.int64) * mb["query_len"]).sum())
    with G.GactEngine(1, -1, -1, -1, tile_size=T, tile_overlap=O, max_tiles=n) as eng:
        eng.upload(G.SET_REF, [mb["ref"].tobytes()])
        eng.upload(G.SET_READS, [mb["query"].tobytes()])
        d = G.make_descs(n)
        for k in ("ref_off", "query_off", "ref_len", "query_len", "reverse", "first"):
            d[k] = mb[k]
        d["ref_set"], d["query_set"] = G.SET_REF, G.SET_READS
        eng.stage(d)
        ms = []
        for _ in range(4):
            eng.run_staged()
            ms.append(eng.last_kernel_ms())
        eng.fetch_staged()
        print(f"T={T:5d} O={O:4d} tiles={n:8d} kernel_ms={min(ms[1:]):8.3f} GCUPS={cells / min(ms[1:]) / 1e6:8.1f} variant={eng.get_kernel()}")
