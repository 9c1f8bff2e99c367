"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
ROOT = os.path.dirname(HERE)


def load_kats():
    return json.load(open(os.path.join(GOLDEN, "align_kats.json")))["vectors"]


def load_random_golden():
    z = np.load(os.path.join(GOLDEN, "align_random.npz"), allow_pickle=True)
    out = []
    for r, q, m, e in zip(z["refs"], z["queries"], z["meta"], z["queues"]):
        out.append(dict(ref=bytes(r), query=bytes(q), scores=tuple(int(x) for x in m[:4]),
                        reverse=int(m[4]), first=int(m[5]), et=int(m[6]), queue=[int(x) for x in e]))
    return out


def oracle_descs(O, mb):
    d = np.zeros(len(mb["ref_off"]), dtype=O.TILE_DESC_DTYPE)
    for k in ("ref_off", "query_off", "ref_len", "query_len", "reverse", "first"):
        d[k] = mb[k]
    return d


def engine_descs(G, mb, ref_set=0, query_set=1):
    d = G.make_descs(len(mb["ref_off"]))
    for k in ("ref_off", "query_off", "ref_len", "query_len", "reverse", "first"):
        d[k] = mb[k]
    d["ref_set"] = ref_set
    d["query_set"] = query_set
    return d


def unpack_all(st_words, n_states, pitch_states):
    """(n, pitch_words) uint32 -> (n, pitch_states) uint8 with zeros past n_states."""
    n = st_words.shape[0]
    k = np.arange(pitch_states)
    vals = (st_words[:, k >> 4] >> (2 * (k & 15)).astype(np.uint32)) & 3
    vals[k[None, :] >= np.asarray(n_states)[:, None]] = 0
    return vals.astype(np.uint8)


def compare_batch(res_gpu, st_gpu, res_cpu, st_cpu):
    """Bit-exact comparison of every field and every state; returns list of mismatching tiles."""
    bad = np.zeros(len(res_gpu), dtype=bool)
    for f in ("score", "max_i", "max_j", "n_states", "i_steps", "j_steps"):
        bad |= res_gpu[f] != res_cpu[f]
    P = min(st_cpu.shape[1], st_gpu.shape[1] * 16)
    g = unpack_all(st_gpu, res_gpu["n_states"], P)
    c = st_cpu[:, :P].copy()
    c[np.arange(P)[None, :] >= res_cpu["n_states"][:, None]] = 0
    bad |= (g != c).any(axis=1)
    return np.nonzero(bad)[0]


def seedtable_cases():
    """Inputs of the seed-position table golden (tests/golden/seedtable_digests.json): (tag, bin-padded reference
    string, k, w, bin_size).  Deterministic, so the generator (make_golden.py, run where /root/reference exists)
    and the tests see the same bytes."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "darwin-gpu_b200"))
    import synth
    rng = np.random.default_rng(2024)
    rnd = lambda n: synth.random_genome(n, rng).tobytes()

    def pad(seqs, b):
        return b"".join(s + b"N" * ((b - len(s) % b) % b) for s in seqs)

    three = [rnd(90000), rnd(40001), rnd(7777)]
    low = rnd(3000) + b"N" * 9000 + rnd(10) + b"A" * 7000 + rnd(2500) + (b"ACGTTGCA" * 700) + rnd(900).lower() + b"RYKM" * 50
    return [
        ("three_k14_w4", pad(three, 64), 14, 4, 64),
        ("three_k12_w8", pad(three, 64), 12, 8, 64),
        ("three_k11_w10_b128", pad(three, 128), 11, 10, 128),
        ("lowcomplexity_k10_w4", pad([low], 64), 10, 4, 64),
        ("lowcomplexity_k8_w1", pad([low], 64), 8, 1, 64),
        ("polyA_k9_w3", b"A" * 20000, 9, 3, 64),
        ("blockedge_k9_w3", pad([rnd(2048 + 3 - 1) + b"C" * 5000 + rnd(100)], 64), 9, 3, 64),
    ]


def table_digest(index, pos):
    import hashlib
    return {"index_entries": int(len(index)), "n_pos": int(len(pos)),
            "index_sha256": hashlib.sha256(np.ascontiguousarray(index, dtype=np.uint32).tobytes()).hexdigest(),
            "pos_sha256": hashlib.sha256(np.ascontiguousarray(pos, dtype=np.uint32).tobytes()).hexdigest()}


TINY_SCHEMES = [(1, -1, -1, -1), (2, -3, -5, -2), (1, -1, -2, -1)]
TINY_ENGINES = [(8, 6), (8, 0)]          # (tile_size, tile_overlap): early_terminate 2 and 8 (>= every length below)


def tiny_tile_batch():
    """Every pair of strings over {A,C,G} of length 1..3, both directions, first and non-first: 6 084 tiles whose DP
    is small enough that all tie situations of align.cpp:138-177 occur (M = I = D, all-non-positive cells, equal
    open/extend, several equal maxima).  Returns a tile batch dict like synth.tile_microbatch()."""
    import itertools
    strings = [("".join(p)).encode() for n in (1, 2, 3) for p in itertools.product("ACG", repeat=n)]
    ref, query, ro, qo, rl, ql, rev, first = [], [], [], [], [], [], [], []
    rpos = qpos = 0
    for r in strings:
        for q in strings:
            for rv in (0, 1):
                for f in (0, 1):
                    ref.append(r); query.append(q)
                    ro.append(rpos); qo.append(qpos); rl.append(len(r)); ql.append(len(q)); rev.append(rv); first.append(f)
                    rpos += len(r); qpos += len(q)
    return dict(ref=np.frombuffer(b"".join(ref), dtype=np.uint8).copy(), query=np.frombuffer(b"".join(query), dtype=np.uint8).copy(),
                ref_off=np.asarray(ro, dtype=np.int64), query_off=np.asarray(qo, dtype=np.int64),
                ref_len=np.asarray(rl, dtype=np.int32), query_len=np.asarray(ql, dtype=np.int32),
                reverse=np.asarray(rev, dtype=np.uint8), first=np.asarray(first, dtype=np.uint8))


def tiny_golden_subset(n):
    """Indices of the tiles whose reference outputs are committed (every 7th tile)."""
    return np.arange(0, n, 7)
