"""CPU-only: the oracle restatement against the reference's golden vectors."""
import numpy as np
import pytest

from helpers import load_kats, load_random_golden


def test_oracle_kats(oracle):
    for v in load_kats():
        got, _ = oracle.align_tile(v["ref"].encode(), v["query"].encode(), tuple(v["scores"]),
                                   v["reverse"], v["first"], v["et"])
        assert got == v["queue"], v


def test_oracle_random_golden(oracle):
    vecs = load_random_golden()
    assert len(vecs) >= 300
    for v in vecs:
        got, _ = oracle.align_tile(v["ref"], v["query"], v["scores"], v["reverse"], v["first"], v["et"])
        assert got == v["queue"]


def test_oracle_vs_reference_live(oracle):
    """Where oracle/_ref was built (build container), compare live on fresh random tiles."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built here")
    import synth
    rng = np.random.default_rng(99)
    for it in range(60):
        T = int(rng.choice([320, 33, 7]))
        rl, ql = int(rng.integers(1, T + 1)), int(rng.integers(1, T + 1))
        g = synth.random_genome(rl, rng)
        q = np.concatenate([synth.error_channel(g, rng)[0], synth.random_genome(T, rng)])[:ql]
        for first in (0, 1):
            a, _ = oracle.align_tile(g.tobytes(), q.tobytes(), (1, -1, -1, -1), it % 2, first, 200)
            b = oracle.ref_align_tile(g.tobytes(), q.tobytes(), (1, -1, -1, -1), it % 2, first, 200)
            assert a == b


def test_oracle_batch_matches_single(oracle):
    import synth
    mb = synth.tile_microbatch(64, seed=5)
    from helpers import oracle_descs
    res, st = oracle.align_batch(mb["ref"], mb["query"], oracle_descs(oracle, mb), n_threads=2)
    for t in range(64):
        r = mb["ref"][mb["ref_off"][t]:mb["ref_off"][t] + mb["ref_len"][t]].tobytes()
        q = mb["query"][mb["query_off"][t]:mb["query_off"][t] + mb["query_len"][t]].tobytes()
        got, rr = oracle.align_tile(r, q, (1, -1, -1, -1), int(mb["reverse"][t]), int(mb["first"][t]), 200)
        assert rr.score == res["score"][t] and rr.n_states == res["n_states"][t]
        assert got[-rr.n_states:] == st[t, :rr.n_states].tolist() or rr.n_states == 0


def test_oracle_properties(oracle):
    """Size-independent properties: empty tiles, all-mismatch tile, state/step accounting."""
    got, r = oracle.align_tile(b"", b"ACGT", first=1)
    assert got == [0, 0, 0]
    got, r = oracle.align_tile(b"ACGT", b"", first=0)
    assert got == [0]
    import synth
    rng = np.random.default_rng(3)
    g = synth.random_genome(320, rng)
    q, _ = synth.error_channel(g, rng)
    got, r = oracle.align_tile(g.tobytes(), q[:320].tobytes(), first=0, et=200)
    st = got[1:]
    assert r.i_steps == sum(s in (3, 2) for s in st) and r.j_steps == sum(s in (3, 1) for s in st)
    assert max(r.i_steps, r.j_steps) <= 200 and len(st) <= 399
