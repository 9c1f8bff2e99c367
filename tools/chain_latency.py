"""Latency anatomy of on-device alignment chains (gact_engine_extend).

  python tools/chain_latency.py [read_kb] [n_copies ...]

One PacBio-like read of read_kb kb (default 30) against a 2 Mbp reference, anchored at its true position near
the read start, extended n_copies times in one call (1 = a lone chain; more = that many identical chains
resident at once, which shows how the per-tile latency grows when chains share an SM sub-partition).
With GACT_LIB=darwin-gpu_b200/libgact_b200_prof.so (make -C darwin-gpu_b200/csrc prof) the chain kernel's
per-phase clocks are printed too (cycles of lane 0 of every warp, summed over warps)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "darwin-gpu_b200"))
import numpy as np
import pygact as G
import synth

kb = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
copies = [int(x) for x in sys.argv[2:]] or [1, 148, 592, 1184, 2368]
rng = np.random.default_rng(5)
genome = synth.random_genome(2000000, rng)
L = int(kb * 1000)
pos = 500000
q, start = synth.error_channel(genome[pos:pos + L], rng)
a = 300                                                # anchor 300 bases into the read
L_ = G.load()
prof = hasattr(L_, "gact_prof_read")
names = ["claim/next-tile", "stage+load_q", "first pass", "dp", "traceback", "consume", "tiles(warp-level)"]
with G.GactEngine(max_tiles=1024) as eng:
    eng.upload(G.SET_REF, [genome.tobytes()])
    eng.upload(G.SET_READS, [q.tobytes()])
    eng.upload(G.SET_READS_RC, [synth.revcomp(q).tobytes()])
    for n in copies:
        calls = np.zeros(n, dtype=G.CALL_DTYPE)
        calls[:] = (0, 0, pos + a, int(start[a]), G.SET_READS, (0, 0, 0))
        times = []
        for it in range(4):
            if prof and it == 3:
                buf = (C.c_ulonglong * 8)()
                L_.gact_prof_read(buf)
            out = eng.extend(calls)
            times.append(eng.last_kernel_ms())
        ms = float(np.median(times[1:]))
        tiles = int(out["n_tiles"][0])
        line = (f"copies {n:5d} read {len(q)} bases tiles/chain {tiles} kernel_ms {ms:.3f} us_per_tile {1e3 * ms / tiles:.2f} "
                f"score {int(out['score'][0])} chain_info {eng.chain_info()}")
        if prof:
            buf = (C.c_ulonglong * 8)()
            L_.gact_prof_read(buf)
            tot = sum(buf[x] for x in range(6))
            wt = max(1, buf[6])
            line += " | cycles per warp-tile: " + ", ".join(f"{names[x]} {buf[x] / wt:.0f}" for x in range(6)) + f" (warp-tiles {buf[6]})"
        print(line, flush=True)
