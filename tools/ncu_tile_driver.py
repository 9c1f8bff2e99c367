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
kw = {}
if "FULL_FRAC" in os.environ:      # e.g. FULL_FRAC=1 FIRST_FRAC=0: only tiles the inter-task kernel takes
    kw = dict(full_frac=float(os.environ["FULL_FRAC"]), first_frac=float(os.environ.get("FIRST_FRAC", "0.055")))
mb = synth.tile_microbatch(n, tile_size=320, seed=42, **kw)
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
    ms = eng.last_kernel_ms()
    res, _ = eng.fetch_staged()
    info = eng.tile_path_info() if hasattr(eng, "tile_path_info") else None
    cells = int((d["ref_len"].astype(np.int64) * d["query_len"]).sum())
    print("kernel_ms", ms, "tiles", n, "gcups", cells / ms / 1e6, "path", info, "checksum", int(res["score"].astype(np.int64).sum()))
