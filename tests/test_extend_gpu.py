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


@pytest.mark.parametrize("tile,overlap,scores", [(320, 120, (1, -1, -1, -1)), (256, 96, (1, -1, -1, -1)),
                                                 (320, 120, (2, -3, -5, -2)), (300, 100, (1, -1, -2, -1)),
                                                 (512, 192, (1, -1, -1, -1)), (1024, 384, (1, -1, -1, -1))])
def test_extend_matches_oracle_gact(pygact, oracle, tile, overlap, scores):
    G, O = pygact, oracle
    genome, reads, rc, calls = make_case(tile + overlap + scores[0])
    with G.GactEngine(*scores, tile_size=tile, tile_overlap=overlap, first_tile_score_threshold=35, max_tiles=64) as eng:
        eng.upload(G.SET_REF, [g.tobytes() for g in genome])
        eng.upload(G.SET_READS, reads)
        eng.upload(G.SET_READS_RC, rc)
        assert eng.extend_supported()
        arr = np.zeros(len(calls), dtype=G.CALL_DTYPE)
        for k, (rs, qs, rp, qp, st) in enumerate(calls):
            arr[k] = (rs, qs, rp, qp, st, (0, 0, 0))
        out = eng.extend(arr)
        ms = eng.last_kernel_ms()
    assert ms > 0
    n_checked = 0
    for k, (rs, qs, rp, qp, st) in enumerate(calls):
        qseq = reads[qs] if st == 1 else rc[qs]
        exp, _ = O.gact_extend(genome[rs].tobytes(), qseq, rp, qp, tile_size=tile, tile_overlap=overlap, thr=35, scores=scores)
        got = out[k]
        assert (got["ab"], got["ae"], got["bb"], got["be"], got["score"], got["first_tile_score"], got["n_tiles"], got["n_cells"]) == \
               (exp.ab, exp.ae, exp.bb, exp.be, exp.score, exp.first_tile_score, exp.n_tiles, exp.n_cells), (k, calls[k])
        n_checked += 1
    assert n_checked > 100 and (out["score"] > 500).sum() > 20


def test_extend_refuses_non_acgt_sets(pygact):
    G = pygact
    with G.GactEngine(max_tiles=16) as eng:
        eng.upload(G.SET_REF, [b"ACGTNACGT" * 50])
        eng.upload(G.SET_READS, [b"ACGT" * 50])
        assert not eng.extend_supported()
        with pytest.raises(G.GactError):
            eng.extend(np.zeros(1, dtype=G.CALL_DTYPE))
