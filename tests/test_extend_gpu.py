"""GPU parity of the on-device candidate extension (gact_engine_extend = the whole GACT() of gact.cpp:48-228)
against the oracle's restatement of GACT(), on PacBio-like reads with random anchors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_case(seed, n_reads=26, genome_len=120000, tile=320):
    import synth
    rng = np.random.default_rng(seed)
    genome = [synth.random_genome(genome_len, rng), synth.random_genome(genome_len // 3, rng)]
    calls = []           # (ref_seq, query_seq, ref_pos, query_pos, query_set)
    reads, rc = [], []
    for r in range(n_reads):
        c = int(rng.integers(0, 2))
        L = int(rng.integers(300, 9000))
        g = genome[c]
        pos = int(rng.integers(0, len(g) - L))
        q, start = synth.error_channel(g[pos:pos + L], rng)
        if len(q) < 50:
            continue
        reads.append(q.tobytes())
        rc.append(synth.revcomp(q).tobytes())
        qi = len(reads) - 1
        # true anchors (a position inside the read and the matching reference position), plus wrong ones
        for _ in range(3):
            a = int(rng.integers(0, L))
            calls.append((c, qi, pos + a, int(start[a]), 1))
        calls.append((c, qi, int(rng.integers(0, len(g))), int(rng.integers(0, len(q))), 1))      # unrelated anchor
        calls.append((1 - c, qi, int(rng.integers(0, len(genome[1 - c]))), int(rng.integers(0, len(q))), 2))
        calls.append((c, qi, 0, 0, 1))                                                            # empty left side
        calls.append((c, qi, len(g), len(q), 1))                                                  # empty right side
    return genome, reads, rc, calls


def _check_against_oracle(O, genome, reads, rc, calls, out, tile, overlap, scores, cache=None):
    n_checked = 0
    for k, (rs, qs, rp, qp, st) in enumerate(calls):
        key = (rs, qs, rp, qp, st)
        if cache is not None and key in cache:
            exp = cache[key]
        else:
            qseq = reads[qs] if st == 1 else rc[qs]
            exp, _ = O.gact_extend(genome[rs].tobytes(), qseq, rp, qp, tile_size=tile, tile_overlap=overlap, thr=35, scores=scores)
            if cache is not None:
                cache[key] = exp
        got = out[k]
        assert (got["ab"], got["ae"], got["bb"], got["be"], got["score"], got["first_tile_score"], got["n_tiles"], got["n_cells"]) == \
               (exp.ab, exp.ae, exp.bb, exp.be, exp.score, exp.first_tile_score, exp.n_tiles, exp.n_cells), (k, calls[k])
        n_checked += 1
    return n_checked


def _calls_array(G, calls):
    arr = np.zeros(len(calls), dtype=G.CALL_DTYPE)
    for k, (rs, qs, rp, qp, st) in enumerate(calls):
        arr[k] = (rs, qs, rp, qp, st, (0, 0, 0))
    return arr


# chain mappings (gact_engine_set_chain_mode): 0 auto, 1 / 2 latency kernel with one / two warps per SM sub-partition,
# 3 throughput kernel, 4 longest chains on the latency kernel + the rest on the throughput kernel
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("tile,overlap,scores", [(320, 120, (1, -1, -1, -1)), (256, 96, (1, -1, -1, -1)),
                                                 (320, 120, (2, -3, -5, -2)), (300, 100, (1, -1, -2, -1)),
                                                 (512, 192, (1, -1, -1, -1)), (1024, 384, (1, -1, -1, -1))])
def test_extend_matches_oracle_gact(pygact, oracle, tile, overlap, scores, mode):
    G, O = pygact, oracle
    if tile > 320 and mode in (1, 2, 4):
        pytest.skip("the latency chain kernel (shared-memory window) exists for tile_size <= 320")
    genome, reads, rc, calls = make_case(tile + overlap + scores[0])
    with G.GactEngine(*scores, tile_size=tile, tile_overlap=overlap, first_tile_score_threshold=35, max_tiles=64) as eng:
        eng.upload(G.SET_REF, [g.tobytes() for g in genome])
        eng.upload(G.SET_READS, reads)
        eng.upload(G.SET_READS_RC, rc)
        assert eng.extend_supported()
        eng.set_chain_mode(mode)
        out = eng.extend(_calls_array(G, calls))
        ms = eng.last_kernel_ms()
        info = eng.chain_info()
    assert ms > 0
    if mode in (1, 2, 3):
        assert info["mode"] == mode
    n_checked = _check_against_oracle(O, genome, reads, rc, calls, out, tile, overlap, scores)
    assert n_checked > 100 and (out["score"] > 500).sum() > 20


@pytest.mark.parametrize("mode", [0, 2, 3, 4])
def test_extend_oversubscribed_queue(pygact, oracle, mode):
    """More calls than chain slots (replicated short and long chains): the queue claims, the dealt first claims, the
    final scan and the long-chain lane all run; every replica must equal the oracle's GACT()."""
    G, O = pygact, oracle
    tile, overlap, scores = 320, 120, (1, -1, -1, -1)
    genome, reads, rc, calls = make_case(99, n_reads=40)
    reps = 6000 // len(calls) + 1
    rng = np.random.default_rng(1)
    big = [calls[i] for i in rng.permutation(len(calls) * reps) % len(calls)]
    with G.GactEngine(*scores, tile_size=tile, tile_overlap=overlap, first_tile_score_threshold=35, max_tiles=64) as eng:
        eng.upload(G.SET_REF, [g.tobytes() for g in genome])
        eng.upload(G.SET_READS, reads)
        eng.upload(G.SET_READS_RC, rc)
        eng.set_chain_mode(mode)
        out = eng.extend(_calls_array(G, big))
        info = eng.chain_info()
    assert len(big) > 5000
    if mode == 4:
        assert info["mode"] == 4 and info["n_long"] >= 8
    assert _check_against_oracle(O, genome, reads, rc, big, out, tile, overlap, scores, cache={}) == len(big)


def test_extend_async_batches_match_sync(pygact):
    """Three batches in flight (each on its own stream with its own buffers) return what the synchronous call returns."""
    G = pygact
    genome, reads, rc, calls = make_case(7, n_reads=30)
    arr = _calls_array(G, calls)
    parts = [arr[0::3], arr[1::3], arr[2::3]]
    with G.GactEngine(first_tile_score_threshold=35, max_tiles=64) as eng:
        eng.upload(G.SET_REF, [g.tobytes() for g in genome])
        eng.upload(G.SET_READS, reads)
        eng.upload(G.SET_READS_RC, rc)
        sync = [eng.extend(p) for p in parts]
        for mode in (0, 3, 1):
            eng.set_chain_mode(mode)
            for p in parts:
                eng.extend_submit(p)
            with pytest.raises(G.GactError):
                eng.extend_submit(parts[0])                    # GACT_MAX_INFLIGHT batches already in flight
            got = [eng.extend_wait() for _ in parts]
            for a, b in zip(sync, got):
                assert (a == b).all()
        with pytest.raises(G.GactError):
            eng.extend_wait()
        eng.extend_submit(arr[:0])                              # empty batch
        assert len(eng.extend_wait()) == 0


def test_extend_with_exceptions_in_the_reference(pygact, oracle):
    """N runs / lower case in the REFERENCE stay on the device chains (an exception row mismatches every ACGT query base,
    which is what raw byte equality gives); a QUERY sequence with such a byte is refused, per sequence."""
    G, O = pygact, oracle
    genome, reads, rc, calls = make_case(11, n_reads=14)
    rng = np.random.default_rng(3)
    g0 = genome[0].copy()
    for _ in range(40):                                    # N runs and lower-case stretches
        p = int(rng.integers(0, len(g0) - 400))
        L = int(rng.integers(1, 300))
        if rng.random() < 0.5:
            g0[p:p + L] = ord("N")
        else:
            g0[p:p + L] |= 0x20
    idx = rng.integers(0, len(g0), size=len(g0) // 50)
    g0[idx] = np.frombuffer(b"NnRY", dtype=np.uint8)[rng.integers(0, 4, size=len(idx))]
    genome = [g0, genome[1]]
    dirty = bytearray(reads[0]); dirty[len(dirty) // 2] = ord("N")
    reads2, rc2 = reads + [bytes(dirty)], rc + [bytes(dirty[::-1])]
    for mode in (0, 1, 3):
        with G.GactEngine(first_tile_score_threshold=35, max_tiles=64) as eng:
            eng.upload(G.SET_REF, [g.tobytes() for g in genome])
            eng.upload(G.SET_READS, reads2)
            eng.upload(G.SET_READS_RC, rc2)
            assert eng.set_bits(G.SET_REF) == 8 and eng.set_bits(G.SET_READS) == 8
            assert eng.extend_supported()
            assert eng.seq_has_exceptions(G.SET_READS, len(reads2) - 1) == 1 and eng.seq_has_exceptions(G.SET_READS, 0) == 0
            assert eng.seq_has_exceptions(G.SET_REF, 0) == 1 and eng.seq_has_exceptions(G.SET_REF, 1) == 0
            eng.set_chain_mode(mode)
            out = eng.extend(_calls_array(G, calls))
            with pytest.raises(G.GactError):
                eng.extend(_calls_array(G, [(0, len(reads2) - 1, 100, 100, 1)]))
        assert _check_against_oracle(O, genome, reads, rc, calls, out, 320, 120, (1, -1, -1, -1)) == len(calls)
    assert (out["score"] > 300).sum() > 5
