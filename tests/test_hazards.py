"""Exactness hazards of SURVEY section 8a (tie order M >= I >= D, ZERO override, >= in the gap flags, last maximum
wins incl. the all-zero tile, early-terminate test before the push, lengths down to 1) on the exhaustive tiny-tile
batch: every pair of strings over {A,C,G} of length 1..3, both directions, first and non-first, three scoring
schemes, early_terminate 2 and 8.  The committed golden (tests/golden/tiny_tiles.npz) holds the queues of the
reference's own AlignWithBT for every 7th of these tiles."""
import os

import numpy as np
import pytest

from helpers import (GOLDEN, TINY_ENGINES, TINY_SCHEMES, compare_batch, engine_descs, oracle_descs, tiny_golden_subset,
                     tiny_tile_batch)


def _golden():
    z = np.load(os.path.join(GOLDEN, "tiny_tiles.npz"))
    return z["flat"].astype(np.int64), z["offs"]


def _queue(res, st, t, first):
    q = [int(res["score"][t])]
    if first:
        q += [int(res["max_i"][t]), int(res["max_j"][t])]
    return q + [int(x) for x in st[t, :res["n_states"][t]]]


def test_tiny_tiles_oracle_matches_reference_golden(oracle):
    mb = tiny_tile_batch()
    n = len(mb["ref_off"])
    assert n == 39 * 39 * 4
    sub = tiny_golden_subset(n)
    flat, offs = _golden()
    k = 0
    seen_states = set()
    for scores in TINY_SCHEMES:
        for T, ov in TINY_ENGINES:
            res, st = oracle.align_batch(mb["ref"], mb["query"], oracle_descs(oracle, mb), scores=scores, et=T - ov,
                                         max_len=T, n_threads=4)
            for t in sub:
                assert _queue(res, st, t, mb["first"][t]) == flat[offs[k]:offs[k + 1]].tolist(), (scores, T, ov, int(t))
                k += 1
            # size-independent invariants on all tiles
            assert (res["n_states"] <= 2 * (T - ov) - 1).all()
            assert (np.maximum(res["i_steps"], res["j_steps"]) <= T - ov).all()
            nf = mb["first"] == 0
            assert (res["max_i"][nf] == mb["ref_len"][nf]).all() and (res["max_j"][nf] == mb["query_len"][nf]).all()
            zero_first = (mb["first"] == 1) & (res["score"] == 0)          # all-zero tile: the last cell is the "maximum"
            assert (res["max_i"][zero_first] == mb["ref_len"][zero_first]).all()
            assert (res["max_j"][zero_first] == mb["query_len"][zero_first]).all()
            for t in range(n):
                seen_states.update(int(x) for x in st[t, :res["n_states"][t]])
    assert k == len(offs) - 1
    assert seen_states == {1, 2, 3}


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [1, 2])
def test_tiny_tiles_cuda_matches_oracle_and_golden(pygact, oracle, variant):
    G = pygact
    mb = tiny_tile_batch()
    n = len(mb["ref_off"])
    sub = tiny_golden_subset(n)
    flat, offs = _golden()
    k = 0
    for scores in TINY_SCHEMES:
        for T, ov in TINY_ENGINES:
            with G.GactEngine(*scores, tile_size=T, tile_overlap=ov, max_tiles=n) as eng:
                eng.set_kernel(variant)
                eng.upload(G.SET_REF, [mb["ref"].tobytes()])
                eng.upload(G.SET_READS, [mb["query"].tobytes()])
                res, st = eng.align_tiles(engine_descs(G, mb))
            ores, ost = oracle.align_batch(mb["ref"], mb["query"], oracle_descs(oracle, mb), scores=scores, et=T - ov,
                                           max_len=T, n_threads=4)
            bad = compare_batch(res, st, ores, ost)
            assert len(bad) == 0, (scores, T, ov, bad[:5])
            from helpers import unpack_all
            u = unpack_all(st, res["n_states"], 2 * (T - ov))
            for t in sub:
                assert _queue(res, u, t, mb["first"][t]) == flat[offs[k]:offs[k + 1]].tolist(), (scores, T, ov, int(t))
                k += 1
