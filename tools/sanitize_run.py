"""Small mixed workload through every kernel family of a bounds-checked build (make -C darwin-gpu_b200/csrc check ->
libgact_b200_check.so: every direction-window and shared-array access range-checked, violations counted per site).
compute-sanitizer is closed on the GPU pool; this is its stand-in.  Prints OK when no check fired and the int32 and the
packed kernels agree.

  GACT_LIB=darwin-gpu_b200/libgact_b200_check.so python tools/sanitize_run.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "darwin-gpu_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pygact as G
import synth

L = G.load()
checked = hasattr(L, "gact_check_read")
SITES = ["window load", "window store", "reference-word read", "traceback base read", "state buffer write", "-", "-", "-"]


def counters():
    if not checked:
        return [0] * 8
    buf = (C.c_ulonglong * 8)()
    assert L.gact_check_read(buf) == 0
    return list(buf)


counters()
n_tiles = 0
for tile, overlap in ((320, 120), (256, 96), (512, 192), (1024, 384), (320, 0), (320, 319), (64, 10)):
    n = 300 if tile >= 512 else 1200
    mb = synth.tile_microbatch(n, tile_size=tile, seed=tile + overlap, full_frac=0.5, first_frac=0.3)
    out = []
    for variant in (1, 2):
        with G.GactEngine(tile_size=tile, tile_overlap=overlap, max_tiles=n) as eng:
            eng.set_kernel(variant)
            eng.upload(G.SET_REF, [mb["ref"].tobytes()])
            eng.upload(G.SET_READS, [mb["query"].tobytes()])
            d = G.make_descs(n)
            for k in ("ref_off", "query_off", "ref_len", "query_len", "reverse", "first"):
                d[k] = mb[k]
            d["ref_set"], d["query_set"] = G.SET_REF, G.SET_READS
            res, st = eng.align_tiles(d)
            out.append(res)
    assert (out[0] == out[1]).all(), f"variants disagree at tile_size {tile}"
    n_tiles += n
# chains in every mapping
from test_extend_gpu import make_case, _calls_array
n_calls = 0
for tile, overlap in ((320, 120), (256, 96), (512, 192)):
    genome, reads, rc, calls = make_case(tile, n_reads=30)
    with G.GactEngine(tile_size=tile, tile_overlap=overlap, max_tiles=64) as eng:
        eng.upload(G.SET_REF, [g.tobytes() for g in genome])
        eng.upload(G.SET_READS, reads)
        eng.upload(G.SET_READS_RC, rc)
        ref = None
        for mode in ((0, 1, 2, 3, 4) if tile <= 320 else (0, 3)):
            eng.set_chain_mode(mode)
            o = eng.extend(_calls_array(G, calls * (30 if mode in (3, 4) else 1)))
            o = o[:len(calls)]
            if ref is None:
                ref = o
            assert (o == ref).all(), f"chain mode {mode} disagrees at tile_size {tile}"
            n_calls += len(calls)
c = counters()
print(f"bounds-checked build: {checked}; {n_tiles} tiles x 2 kernel variants, {n_calls} chain calls; violations per site: "
      + ", ".join(f"{s} {v}" for s, v in zip(SITES[:5], c[:5])))
assert sum(c) == 0, "range check fired"
print("OK")
