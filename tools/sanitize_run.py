"""Small mixed batch through every kernel variant (for compute-sanitizer memcheck): prints OK when results match
between the int32 and the packed kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "darwin-gpu_b200"))
import pygact as G
import synth

for tile, overlap in ((320, 120), (256, 96), (512, 192)):
    mb = synth.tile_microbatch(600, tile_size=tile, seed=tile, full_frac=0.5, first_frac=0.3)
    out = []
    for variant in (1, 2):
        with G.GactEngine(tile_size=tile, tile_overlap=overlap, max_tiles=600) as eng:
            eng.set_kernel(variant)
            eng.upload(G.SET_REF, [mb["ref"].tobytes()])
            eng.upload(G.SET_READS, [mb["query"].tobytes()])
            d = G.make_descs(600)
            for k in ("ref_off", "query_off", "ref_len", "query_len", "reverse", "first"):
                d[k] = mb[k]
            d["ref_set"], d["query_set"] = G.SET_REF, G.SET_READS
            res, st = eng.align_tiles(d)
            out.append(res)
    assert (out[0] == out[1]).all(), f"variants disagree at tile_size {tile}"
print("OK")
