"""Where does the e2e leg of bench.py spend host time?  Per-call timing of submit/wait over one config-2 pass."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "darwin-gpu_b200"))
import numpy as np
import pygact as G
import bench

n = 1 << 20
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
mb = bench.make_batch(n, 42)
cells = int((mb["ref_len"].astype(np.int64) * mb["query_len"]).sum())
eng = G.GactEngine(*bench.SCORES, tile_size=bench.TILE, tile_overlap=bench.OVERLAP, device=0, max_tiles=chunk)
eng.upload(G.SET_REF, [mb["ref"].tobytes()]); eng.upload(G.SET_READS, [mb["query"].tobytes()])
d = G.make_descs(n)
for k in ("ref_off", "query_off", "ref_len", "query_len", "reverse", "first"):
    d[k] = mb[k]
d["ref_set"], d["query_set"] = G.SET_REF, G.SET_READS
res = np.zeros(n, dtype=G.TILE_RESULT_DTYPE); st = np.zeros((n, eng.pitch), dtype=np.uint32)
bounds = [(lo, min(lo + chunk, n)) for lo in range(0, n, chunk)]
for rep in range(3):
    ts, tw = [], []
    pend = []
    k0 = eng.stats()["kernel_ms"]
    t00 = time.perf_counter()
    for lo, hi in bounds:
        t0 = time.perf_counter(); eng.submit(d[lo:hi]); ts.append(time.perf_counter() - t0)
        pend.append((lo, hi))
        if len(pend) == G.MAX_INFLIGHT:
            a, b = pend.pop(0)
            t0 = time.perf_counter(); eng.wait(res[a:b], st[a:b]); tw.append(time.perf_counter() - t0)
    while pend:
        a, b = pend.pop(0)
        t0 = time.perf_counter(); eng.wait(res[a:b], st[a:b]); tw.append(time.perf_counter() - t0)
    tot = time.perf_counter() - t00
    print(f"pass {rep}: total {tot*1e3:.2f} ms  ({cells/tot/1e9:.0f} GCUPS)  sum submit {sum(ts)*1e3:.2f}  sum wait {sum(tw)*1e3:.2f}  "
          f"kernel_ms sum {eng.stats()['kernel_ms']-k0:.2f}")
    print("  submit ms:", " ".join(f"{x*1e3:.2f}" for x in ts))
    print("  wait   ms:", " ".join(f"{x*1e3:.2f}" for x in tw))
