"""GPU parity of the device-side D-SOFT filter: candidate streams (per query, in emission order) against the
reference's golden candidates (tests/golden/dsoft_e2e_small.npz, produced by the unmodified
SeedPosTable::DSOFT) and against the host implementation on fresh random data."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def host_lib():
    H = C.CDLL(os.path.join(ROOT, "darwin-gpu_b200", "libdarwin_host.so"))
    H.dh_seed_table_new.restype = C.c_void_p
    H.dh_seed_table_new.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    H.dh_seed_table_free.argtypes = [C.c_void_p]
    H.dh_dsoft.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    H.dh_seed_table_arrays.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    return H


def table_arrays(H, t):
    ip, ie, pp, npos, mo = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64(), C.c_uint32()
    H.dh_seed_table_arrays(t, C.byref(ip), C.byref(ie), C.byref(pp), C.byref(npos), C.byref(mo))
    return ip, ie.value, pp, npos.value, mo.value


def run_case(G, H, refs, reads, k=14, w=4, bin_size=64, num_seeds=800, threshold=21, expect=None):
    import synth
    refstr = b"".join(s + b"N" * ((bin_size - len(s) % bin_size) % bin_size) for s in refs)
    t = H.dh_seed_table_new(refstr, len(refstr), k, 32, bin_size, w, 4)
    assert t
    ip, ie, pp, npos, mo = table_arrays(H, t)
    rc_reads = [synth.revcomp(np.frombuffer(r, dtype=np.uint8)).tobytes() for r in reads]
    with G.GactEngine(max_tiles=16) as eng:
        eng.upload(G.SET_READS, reads)
        eng.upload(G.SET_READS_RC, rc_reads)
        ds = G.Dsoft(eng, ip, ie, pp, npos, kmer_size=k, window_size=w, bin_size=bin_size, max_occ=mo,
                     num_seeds=num_seeds, threshold=threshold)
        sets, idx = [], []
        for i in range(len(reads)):                     # per read: forward strand, then reverse complement
            sets += [G.SET_READS, G.SET_READS_RC]
            idx += [i, i]
        got = ds.run(sets, idx, cap=64)                 # small capacity: exercises the grow-and-retry path
        ms = ds.last_kernel_ms()
        # asynchronous form: two batches in flight (halves of the query list), same candidates
        half = (len(sets) // 4) * 2
        ds.submit(sets[:half], idx[:half], cap=max(64, len(got)))
        ds.submit(sets[half:], idx[half:], cap=max(64, len(got)))
        a, b = ds.wait(), ds.wait()
        b = b.copy()
        b["query"] += half
        both = np.concatenate([a, b])
        assert len(both) == len(got) and (both == got).all()
        ds.close()
    # host implementation (pinned to the reference by tests/test_host.py)
    n_q = len(sets)
    for qi in range(n_q):
        strand = reads[qi // 2] if qi % 2 == 0 else rc_reads[qi // 2]
        buf = np.zeros(1 << 14, dtype=np.uint64)
        n = H.dh_dsoft(t, strand, len(strand), num_seeds, threshold, 1000000, 2500000, buf.ctypes.data, 1 << 14)
        mine = got[got["query"] == qi]
        assert len(mine) == n, (qi, len(mine), n)
        assert (mine["seq"] == np.arange(n)).all()
        packed = (mine["hit"].astype(np.uint64) << np.uint64(32)) | mine["offset"].astype(np.uint64)
        assert (packed == buf[:n]).all(), qi
        if expect is not None:
            assert n == expect["counts"][qi]
    H.dh_seed_table_free(t)
    return got, ms


def read_fasta_simple(path):
    recs, name, cur = [], None, []
    for ln in open(path, "rb").read().split(b"\n"):
        if ln.startswith(b">"):
            if name is not None:
                recs.append(b"".join(cur))
            name, cur = ln, []
        elif ln:
            cur.append(ln)
    if name is not None:
        recs.append(b"".join(cur))
    return recs


def test_dsoft_gpu_matches_reference_golden(pygact):
    H = host_lib()
    z = np.load(os.path.join(GOLD, "dsoft_e2e_small.npz"))
    refs = read_fasta_simple(os.path.join(GOLD, "e2e_small", "ref.fasta"))
    reads = read_fasta_simple(os.path.join(GOLD, "e2e_small", "reads.fasta"))
    got, ms = run_case(pygact, H, refs, reads, expect={"counts": z["counts"].tolist()})
    packed = (got["hit"].astype(np.uint64) << np.uint64(32)) | got["offset"].astype(np.uint64)
    assert (packed == z["cands"]).all()                 # same candidates, same order as the reference's DSOFT
    assert ms > 0


@pytest.mark.parametrize("k,w,bin_size,num_seeds,threshold", [(14, 4, 64, 800, 21), (12, 3, 128, 50, 24), (15, 7, 32, 2000, 15)])
def test_dsoft_gpu_matches_host_on_random_data(pygact, k, w, bin_size, num_seeds, threshold):
    import synth
    H = host_lib()
    rng = np.random.default_rng(k * 100 + w)
    genome = [synth.random_genome(220000, rng), synth.random_genome(70000, rng)]
    # a repeat-rich piece: many hits per seed, several hits per bin
    rep = synth.random_genome(500, rng)
    genome.append(np.concatenate([rep] * 60 + [synth.random_genome(3000, rng)]))
    _, reads = synth.sample_reads(genome, 400000, rng, mean=4000, sd=2500, lo=20, hi=12000)
    reads = [r.tobytes() for r in reads] + [b"ACGT" * 5, b"A" * 17, np.tile(rep, 8).tobytes()]
    run_case(pygact, H, [g.tobytes() for g in genome], reads, k=k, w=w, bin_size=bin_size,
             num_seeds=num_seeds, threshold=threshold)
