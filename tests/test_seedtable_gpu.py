"""GPU parity of the device-side seed-position table builder (gact_seed_table_build) against the host builder
(host/seed_table.cpp, itself pinned to the reference's SeedPosTable by tests/test_host.py): index_table_ and
pos_table_ must be identical word for word, and D-SOFT on the borrowed device table must return the same
candidates as on the uploaded host table."""
import ctypes as C
import os

import numpy as np
import pytest

from test_dsoft_gpu import host_lib, table_arrays

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def host_tables(H, refstr, k, w, bin_size, occ=32):
    t = H.dh_seed_table_new(refstr, len(refstr), k, occ, bin_size, w, 4)
    assert t
    ip, ie, pp, npos, mo = table_arrays(H, t)
    index = np.ctypeslib.as_array((C.c_uint32 * ie).from_address(ip.value)).copy()
    pos = np.ctypeslib.as_array((C.c_uint32 * max(npos, 1)).from_address(pp.value)).copy()[:npos]
    H.dh_seed_table_free(t)
    return index, pos, mo


def check(G, H, eng, refstr, k, w, bin_size=64):
    index, pos, mo = host_tables(H, refstr, k, w, bin_size)
    tab = G.SeedTable(eng, refstr, kmer_size=k, seed_occurence_multiple=32, bin_size=bin_size, window_size=w)
    try:
        assert tab.index_entries == len(index) == 4 ** k + 1
        assert tab.num_minimizers == len(pos), (tab.num_minimizers, len(pos))
        assert tab.max_occ == mo
        gi, gp = tab.download()
        assert np.array_equal(gp, pos)
        assert np.array_equal(gi, index)
        assert tab.build_ms > 0
    finally:
        tab.close()
    return len(pos)


def pad(seqs, bin_size=64):
    return b"".join(s + b"N" * ((bin_size - len(s) % bin_size) % bin_size) for s in seqs)


def test_seed_table_matches_host_builder(pygact):
    G = pygact
    import synth
    H = host_lib()
    rng = np.random.default_rng(11)
    genome = [synth.random_genome(n, rng).tobytes() for n in (150000, 70001, 33333)]
    with G.GactEngine(max_tiles=16) as eng:
        n = check(G, H, eng, pad(genome), 14, 4)
        assert n > 50000
        check(G, H, eng, pad(genome), 12, 8)
        check(G, H, eng, pad(genome), 13, 12)
        check(G, H, eng, pad(genome, 128), 11, 10, bin_size=128)
        check(G, H, eng, pad(genome), 8, 1)           # w = 1: every position is its own minimizer


def test_seed_table_matches_reference_golden(pygact):
    """Tables built on the device against the digests of the reference's own SeedPosTable constructor
    (tests/golden/seedtable_digests.json, written by make_golden.py where /root/reference exists)."""
    import json
    from helpers import seedtable_cases, table_digest
    G = pygact
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "seedtable_digests.json")))
    with G.GactEngine(max_tiles=16) as eng:
        for tag, refstr, k, w, b in seedtable_cases():
            tab = G.SeedTable(eng, refstr, kmer_size=k, seed_occurence_multiple=32, bin_size=b, window_size=w)
            try:
                index, pos = tab.download()
            finally:
                tab.close()
            d = table_digest(index, pos)
            for key in ("index_entries", "n_pos", "index_sha256", "pos_sha256"):
                assert d[key] == gold[tag][key], (tag, key)


def test_seed_table_low_complexity_and_block_edges(pygact):
    """Constant-minimum runs (poly-A, N padding, tandem repeats) longer than a thread block's 2048 positions carry
    their run start across blocks; lower-case and non-ACGT bytes follow ntcoding.cpp:60-72."""
    G = pygact
    import synth
    H = host_lib()
    rng = np.random.default_rng(12)
    rnd = lambda n: synth.random_genome(n, rng).tobytes()
    cases = [
        b"A" * 20000,                                               # the all-zero start: last_m = 0 is never "changed"
        rnd(3000) + b"N" * 9000 + rnd(10) + b"A" * 7000 + rnd(2500),
        rnd(2048 + 3 - 1) + b"C" * 5000 + rnd(100),
        (b"ACGTTGCA" * 4000) + rnd(777) + (b"AC" * 3000),
        rnd(5000).lower() + b"RYKM" * 300 + rnd(5000),
        rnd(2047), rnd(2048), rnd(2049), rnd(2051), rnd(2052), rnd(4099), rnd(4100),
        rnd(17), rnd(18), rnd(19), rnd(31), rnd(32), rnd(33), b"ACGT", b"",
        b"T" * 2051 + b"G" * 2051 + b"T" * 2051,
    ]
    with G.GactEngine(max_tiles=16) as eng:
        for k, w in ((14, 4), (9, 3), (6, 5)):
            for s in cases:
                check(G, H, eng, s, k, w)


def test_dsoft_on_device_built_table(pygact):
    G = pygact
    import synth
    H = host_lib()
    rng = np.random.default_rng(13)
    genome = [synth.random_genome(200000, rng) for _ in range(2)]
    names, reads = synth.sample_reads(genome, 300000, rng, mean=3000.0, sd=1000.0, lo=500, hi=8000)
    refstr = pad([g.tobytes() for g in genome])
    rc = [synth.revcomp(r) for r in reads]
    t = H.dh_seed_table_new(refstr, len(refstr), 14, 32, 64, 4, 4)
    ip, ie, pp, npos, mo = table_arrays(H, t)
    with G.GactEngine(max_tiles=16) as eng:
        eng.upload(G.SET_READS, [r.tobytes() for r in reads])
        eng.upload(G.SET_READS_RC, [r.tobytes() for r in rc])
        sets, idx = [], []
        for i in range(len(reads)):
            sets += [G.SET_READS, G.SET_READS_RC]
            idx += [i, i]
        ds_host = G.Dsoft(eng, ip, ie, pp, npos, max_occ=mo)
        a = ds_host.run(sets, idx)
        ds_host.close()
        tab = G.SeedTable(eng, refstr)
        ds_dev = G.Dsoft(eng, table=tab)
        b = ds_dev.run(sets, idx)
        ds_dev.close()
        tab.close()
    H.dh_seed_table_free(t)
    assert len(a) > 50
    key = lambda x: np.lexsort((x["seq"], x["query"]))
    assert np.array_equal(a[key(a)], b[key(b)])
