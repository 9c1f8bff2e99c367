"""Small stand-alone driver for profiling the on-device chain kernel and the D-SOFT kernel:
2 Mbp reference, ~3 MB of PacBio-like reads; D-SOFT on the device, then gact_engine_extend on its candidates."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "darwin-gpu_b200"))
import numpy as np
import pygact as G
import synth

H = C.CDLL(os.path.join(ROOT, "darwin-gpu_b200", "libdarwin_host.so"))
H.dh_seed_table_new.restype = C.c_void_p
H.dh_seed_table_new.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
H.dh_seed_table_arrays.argtypes = [C.c_void_p] + [C.c_void_p] * 5

rng = np.random.default_rng(1)
genome = [synth.random_genome(1000000, rng) for _ in range(2)]
names, reads = synth.sample_reads(genome, int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 3000000, rng)
rc = [synth.revcomp(r) for r in reads]
refstr = b"".join(g.tobytes() for g in genome)           # 1 Mbp pieces are multiples of the bin size
t = H.dh_seed_table_new(refstr, len(refstr), 14, 32, 64, 4, 8)
ip, ie, pp, npos, mo = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64(), C.c_uint32()
H.dh_seed_table_arrays(t, C.byref(ip), C.byref(ie), C.byref(pp), C.byref(npos), C.byref(mo))
with G.GactEngine(max_tiles=1024) as eng:
    eng.upload(G.SET_REF, [g.tobytes() for g in genome])
    eng.upload(G.SET_READS, [r.tobytes() for r in reads])
    eng.upload(G.SET_READS_RC, [r.tobytes() for r in rc])
    ds = G.Dsoft(eng, ip, ie.value, pp, npos.value, max_occ=mo.value)
    sets, idx = [], []
    for i in range(len(reads)):
        sets += [G.SET_READS, G.SET_READS_RC]
        idx += [i, i]
    for _ in range(2):
        cands = ds.run(sets, idx)
    dsoft_ms = ds.last_kernel_ms()
    calls = np.zeros(len(cands), dtype=G.CALL_DTYPE)
    for k, c in enumerate(cands):
        hit = int(c["hit"])
        chrom = hit // 1000000
        calls[k] = (chrom, int(c["query"]) // 2, hit - chrom * 1000000, int(c["offset"]),
                    G.SET_READS if c["query"] % 2 == 0 else G.SET_READS_RC, (0, 0, 0))
    times = []
    for _ in range(7):
        out = eng.extend(calls)
        times.append(eng.last_kernel_ms())
    info = eng.chain_info()
    chain_ms = float(np.median(times[1:]))
    chain_min = min(times[1:])
    ds.close()
cells = int(out["n_cells"].sum())
print(f"chain_info {info} reads {len(reads)} strand-queries {len(sets)} candidates {len(cands)} dsoft_kernel_ms {dsoft_ms:.3f} "
      f"chain_kernel_ms {chain_ms:.3f} (min {chain_min:.3f}) tiles {int(out['n_tiles'].sum())} cells {cells} chain_gcups {cells / chain_ms / 1e6:.1f}")
