"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


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
