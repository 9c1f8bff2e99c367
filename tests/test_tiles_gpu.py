"""GPU parity tests: the CUDA tile kernels, called through the C ABI, against the oracle and
the reference's golden vectors.  Bit-exact on score, max position, step counts and every
traceback state."""
import numpy as np
import pytest

from helpers import compare_batch, engine_descs, load_kats, load_random_golden, oracle_descs

pytestmark = pytest.mark.gpu

VARIANTS = [1, 2]        # 1 = int32 DPX kernel, 2 = packed s16x2 DPX kernel


def _engine(G, variant, soft=False, **kw):
    """soft: return None instead of skipping the whole test when the variant cannot run these parameters."""
    eng = G.GactEngine(**kw)
    try:
        eng.set_kernel(variant)
    except G.GactError:
        eng.close()
        if soft:
            return None
        pytest.skip(f"kernel variant {variant} not available for these parameters")
    return eng


@pytest.mark.parametrize("variant", VARIANTS)
def test_golden_kats(pygact, variant):
    G = pygact
    for v in load_kats():
        ma, mi, go, ge = v["scores"]
        T = 320
        with _engine(G, variant, match=ma, mismatch=mi, gap_open=go, gap_extend=ge,
                     tile_size=T, tile_overlap=T - v["et"], max_tiles=8) as eng:
            q = G.align_with_bt(eng, v["ref"].encode(), v["query"].encode(), v["reverse"], v["first"])
        assert q == v["queue"], (v, q)


@pytest.mark.parametrize("variant", VARIANTS)
def test_golden_random_tiles(pygact, variant):
    """The 360 tiles whose expected queues came from the reference's own AlignWithBT."""
    G = pygact
    vecs = load_random_golden()
    groups = {}
    for v in vecs:
        groups.setdefault((v["scores"], v["et"]), []).append(v)
    ran = 0
    for (sc, et), vs in groups.items():
        T = max(320, et + 1)
        eng = _engine(G, variant, soft=True, match=sc[0], mismatch=sc[1], gap_open=sc[2], gap_extend=sc[3],
                      tile_size=T, tile_overlap=T - et, max_tiles=256)
        if eng is None:
            continue        # outside the packed kernel's 16-bit score range: only this group is left to the int32 kernel
        with eng:
            out = G.align_batch(eng, [v["ref"] for v in vs], [v["query"] for v in vs],
                                [v["reverse"] for v in vs], [v["first"] for v in vs])
        for v, q in zip(vs, out):
            assert q == v["queue"], (sc, et, len(v["ref"]), len(v["query"]), v["reverse"], v["first"])
        ran += len(vs)
    assert ran >= (len(vecs) if variant == 1 else 340), f"only {ran} of {len(vecs)} golden tiles ran on variant {variant}"


def _run_microbatch(G, O, variant, n, seed, tile=320, overlap=120, scores=(1, -1, -1, -1), **mbkw):
    import synth
    mb = synth.tile_microbatch(n, tile_size=tile, seed=seed, **mbkw)
    with _engine(G, variant, match=scores[0], mismatch=scores[1], gap_open=scores[2], gap_extend=scores[3],
                 tile_size=tile, tile_overlap=overlap, max_tiles=n) as eng:
        eng.upload(G.SET_REF, [mb["ref"].tobytes()])
        eng.upload(G.SET_READS, [mb["query"].tobytes()])
        assert eng.set_bits(G.SET_REF) == 2 and eng.set_bits(G.SET_READS) == 2
        res, st = eng.align_tiles(engine_descs(G, mb))
    ores, ost = O.align_batch(mb["ref"], mb["query"], oracle_descs(O, mb), scores=scores,
                              et=tile - overlap, max_len=tile, n_threads=8)
    bad = compare_batch(res, st, ores, ost)
    assert len(bad) == 0, f"{len(bad)} of {n} tiles differ, first {bad[:5]}: gpu {res[bad[:3]]} cpu {ores[bad[:3]]}"
    return res


@pytest.mark.parametrize("variant", VARIANTS)
def test_microbatch_default_params(pygact, oracle, variant):
    """Config-2 shaped tiles (PacBio-like 15 % error, 18 % ragged edge tiles, 5.5 % first)."""
    res = _run_microbatch(pygact, oracle, variant, 4096, seed=42)
    assert res["score"].max() > 100


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("scores", [(2, -3, -5, -2), (1, -1, -2, -1), (5, -4, -10, -1), (1, -3, 0, 0), (3, -1, -1, -4)])
def test_microbatch_scoring_schemes(pygact, oracle, variant, scores):
    _run_microbatch(pygact, oracle, variant, 1024, seed=7, scores=scores, first_frac=0.3)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("tile,overlap", [(256, 96), (512, 192), (1024, 384), (320, 0), (320, 319), (64, 10), (100, 37)])
def test_tile_size_sweep(pygact, oracle, variant, tile, overlap):
    """Config-5 tile sizes plus odd shapes (early_terminate = 1, no overlap)."""
    n = 256 if tile >= 512 else 768
    _run_microbatch(pygact, oracle, variant, n, seed=tile + overlap, tile=tile, overlap=overlap, first_frac=0.2)


@pytest.mark.parametrize("variant", VARIANTS)
def test_all_edge_and_all_first(pygact, oracle, variant):
    _run_microbatch(pygact, oracle, variant, 1500, seed=3, full_frac=0.0, first_frac=1.0)
    _run_microbatch(pygact, oracle, variant, 1500, seed=4, full_frac=0.0, first_frac=0.0)


@pytest.mark.parametrize("variant", VARIANTS)
def test_unrelated_sequences_zero_states(pygact, oracle, variant):
    """Random pairs: low scores, many ZERO stops, last-maximum rule on ties."""
    _run_microbatch(pygact, oracle, variant, 1024, seed=9, err=(0.5, 0.2, 0.2), first_frac=0.5)


@pytest.mark.parametrize("variant", VARIANTS)
def test_raw_byte_semantics_8bit_sets(pygact, oracle, variant):
    """N / lower-case bases: sets fall back to 8 bits per base; 'N'=='N' matches, 'a'!='A'."""
    G, O = pygact, oracle
    import synth
    mb = synth.tile_microbatch(512, seed=21, first_frac=0.2)
    rng = np.random.default_rng(1)
    ref, qry = mb["ref"].copy(), mb["query"].copy()
    for buf in (ref, qry):
        idx = rng.integers(0, len(buf), size=len(buf) // 20)
        buf[idx] = np.frombuffer(b"Nacgtn", dtype=np.uint8)[rng.integers(0, 6, size=len(idx))]
    mb["ref"], mb["query"] = ref, qry
    with _engine(G, variant, max_tiles=512) as eng:
        eng.upload(G.SET_REF, [ref.tobytes()])
        eng.upload(G.SET_READS, [qry.tobytes()])
        assert eng.set_bits(G.SET_REF) == 8 and eng.set_bits(G.SET_READS) == 8
        res, st = eng.align_tiles(engine_descs(G, mb))
    ores, ost = O.align_batch(ref, qry, oracle_descs(O, mb), n_threads=8)
    assert len(compare_batch(res, st, ores, ost)) == 0


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("where", ["ref", "query_some", "both_some"])
def test_exceptions_routed_per_tile(pygact, oracle, variant, where):
    """Bytes other than ACGT in the reference only (score-table kernels with sentinel rows), in some query windows (those
    tiles go to the raw-byte kernels, the rest of the batch stays on the table kernels) or in both: raw byte equality
    (align.cpp:134) on every tile."""
    G, O = pygact, oracle
    import synth
    mb = synth.tile_microbatch(1500, seed=31, first_frac=0.25)
    rng = np.random.default_rng(5)
    ref, qry = mb["ref"].copy(), mb["query"].copy()
    sub = np.frombuffer(b"NNNacgtnRY", dtype=np.uint8)

    def sprinkle(buf, frac, runs):
        idx = rng.integers(0, len(buf), size=int(len(buf) * frac))
        buf[idx] = sub[rng.integers(0, len(sub), size=len(idx))]
        for _ in range(runs):
            p = int(rng.integers(0, len(buf) - 200))
            buf[p:p + int(rng.integers(1, 150))] = ord("N")

    if where in ("ref", "both_some"):
        sprinkle(ref, 0.01, 30)
    if where in ("query_some", "both_some"):
        sprinkle(qry[: len(qry) // 3], 0.002, 10)        # only the first third of the query buffer: most windows stay clean
    mb["ref"], mb["query"] = ref, qry
    with _engine(G, variant, max_tiles=1500) as eng:
        eng.upload(G.SET_REF, [ref.tobytes()])
        eng.upload(G.SET_READS, [qry.tobytes()])
        assert eng.set_bits(G.SET_REF) == (8 if where != "query_some" else 2)
        res, st = eng.align_tiles(engine_descs(G, mb))
    ores, ost = O.align_batch(ref, qry, oracle_descs(O, mb), n_threads=8)
    bad = compare_batch(res, st, ores, ost)
    assert len(bad) == 0, f"{len(bad)} tiles differ, first {bad[:5]}"


@pytest.mark.parametrize("variant", VARIANTS)
def test_empty_and_tiny_tiles(pygact, variant):
    G = pygact
    with _engine(G, variant, max_tiles=16) as eng:
        out = G.align_batch(eng, [b"A", b"A", b"ACGT", b"C", b"ACGT"], [b"A", b"C", b"A", b"ACGT", b"ACGT"],
                            [0, 0, 1, 1, 0], [1, 1, 0, 1, 0])
    assert out[0] == [1, 1, 1, 3]
    assert out[1] == [0, 1, 1]
    assert out[4] == [4, 3, 3, 3, 3]
    with _engine(G, variant, max_tiles=16) as eng:
        eng.upload(G.SET_REF, [b"ACGT"])
        eng.upload(G.SET_READS, [b"ACGT"])
        d = G.make_descs(3)
        d["ref_len"] = [0, 4, 0]
        d["query_len"] = [4, 0, 0]
        d["query_set"] = G.SET_READS
        d["first"] = [1, 0, 1]
        res, st = eng.align_tiles(d)
        assert res["score"].tolist() == [0, 0, 0] and res["n_states"].tolist() == [0, 0, 0]
        assert res["max_i"].tolist() == [0, 4, 0] and res["max_j"].tolist() == [0, 0, 0]
        res0, _ = eng.align_tiles(G.make_descs(0))
        assert len(res0) == 0


@pytest.mark.parametrize("band", ["default", "5"])
@pytest.mark.parametrize("tile,overlap,scores", [(320, 120, (1, -1, -1, -1)), (256, 96, (2, -3, -5, -2)), (320, 0, (1, -1, -2, -1)),
                                                 (512, 192, (1, -1, -1, -1)), (64, 10, (1, -3, 0, 0)), (1024, 384, (1, -1, -1, -1))])
def test_inter_task_kernel_matches_oracle(pygact, oracle, monkeypatch, band, tile, overlap, scores):
    """Full, non-first tiles of a batch on the inter-task kernel (one lane per pair of tiles, direction codes only for a band
    around the diagonal); with a 5-wide band most tracebacks leave it and the tiles are handed back to the wavefront kernel.
    Every tile of the mixed batch (full / edge / first) must equal the oracle either way.  (A reference set with bytes other
    than ACGT keeps all its tiles on the wavefront kernels: test_exceptions_routed_per_tile.)"""
    G, O = pygact, oracle
    import synth
    monkeypatch.setenv("GACT_IT_MIN", "64")
    if band != "default":
        monkeypatch.setenv("GACT_IT_BAND", band)          # default: max(32, early_terminate / 8)
    n = 700 if tile >= 512 else 2600
    mb = synth.tile_microbatch(n, tile_size=tile, seed=tile + len(band), full_frac=0.9, first_frac=0.1)
    ref = mb["ref"]
    with G.GactEngine(*scores, tile_size=tile, tile_overlap=overlap, max_tiles=n) as eng:
        eng.upload(G.SET_REF, [ref.tobytes()])
        eng.upload(G.SET_READS, [mb["query"].tobytes()])
        res, st = eng.align_tiles(engine_descs(G, mb))
        info = eng.tile_path_info()
    eligible = int(((mb["ref_len"] == tile) & (mb["query_len"] == tile) & (mb["first"] == 0)).sum())
    assert info["inter_task"] == (eligible // 64) * 64 and info["inter_task"] > 0.5 * n
    if band == "5":
        assert info["handed_back"] > 0                             # a 15 % error channel wanders more than 5 off the diagonal
    elif scores == (1, -1, -1, -1):
        assert info["handed_back"] < 0.05 * info["inter_task"]
    ores, ost = O.align_batch(ref, mb["query"], oracle_descs(O, mb), scores=scores, et=tile - overlap, max_len=tile, n_threads=8)
    bad = compare_batch(res, st, ores, ost)
    assert len(bad) == 0, f"{len(bad)} of {n} tiles differ, first {bad[:5]}: gpu {res[bad[:3]]} cpu {ores[bad[:3]]}"


@pytest.mark.parametrize("tile,overlap,scores", [(320, 120, (1, -1, -1, -1)), (256, 96, (2, -3, -5, -2)), (320, 0, (1, -1, -2, -1)),
                                                 (300, 100, (1, -1, -1, -1)), (64, 10, (1, -3, 0, 0)), (200, 190, (1, -2, -3, -1))])
def test_narrow_mapping_matches_oracle(pygact, oracle, monkeypatch, tile, overlap, scores):
    """Ragged batch (every window length uniform in 1..tile_size): non-first tiles whose query window fits strips of half the
    width run on the narrow mapping of the wavefront kernel (reference windows up to the full tile size), the others on the
    regular one.  Every tile must equal the oracle, and the results must not depend on the mapping (GACT_NARROW=0)."""
    G, O = pygact, oracle
    import synth
    n = 3000
    mb = synth.tile_microbatch(n, tile_size=tile, seed=3 * tile + overlap, full_frac=0.1, first_frac=0.1)
    # some tiles at the corners of the narrow mapping's range: widest query it takes, with the longest reference
    cols = (4 if tile <= 256 else 5) * 32
    k = np.flatnonzero(mb["first"] == 0)[:40]
    mb["query_len"][k[:20]] = min(cols, tile)
    mb["ref_len"][k[:20]] = tile
    mb["query_len"][k[20:]] = min(cols + 1, tile)              # one column too wide: regular mapping
    mb["ref_off"] = np.clip(mb["ref_off"], 0, len(mb["ref"]) - tile - 1)
    ref = mb["ref"]
    out = []
    for narrow in ("1", "0"):
        monkeypatch.setenv("GACT_NARROW", narrow)
        with G.GactEngine(*scores, tile_size=tile, tile_overlap=overlap, max_tiles=n) as eng:
            eng.upload(G.SET_REF, [ref.tobytes()])
            eng.upload(G.SET_READS, [mb["query"].tobytes()])
            out.append(eng.align_tiles(engine_descs(G, mb)) + (eng.stats()["kernel_launches"],))
    assert out[0][2] == out[1][2] + 1                               # the narrow group had its own launch
    assert (out[0][0] == out[1][0]).all()
    from helpers import unpack_all                                  # state words beyond n_states are not defined
    P = out[0][1].shape[1] * 16
    assert (unpack_all(out[0][1], out[0][0]["n_states"], P) == unpack_all(out[1][1], out[1][0]["n_states"], P)).all()
    ores, ost = O.align_batch(ref, mb["query"], oracle_descs(O, mb), scores=scores, et=tile - overlap, max_len=tile, n_threads=8)
    bad = compare_batch(out[0][0], out[0][1], ores, ost)
    assert len(bad) == 0, f"{len(bad)} of {n} tiles differ, first {bad[:5]}: gpu {out[0][0][bad[:3]]} cpu {ores[bad[:3]]}"


def test_async_submit_wait_matches_sync(pygact):
    G = pygact
    import synth
    mb = synth.tile_microbatch(3000, seed=77)
    with G.GactEngine(max_tiles=1024) as eng:
        eng.upload(G.SET_REF, [mb["ref"].tobytes()])
        eng.upload(G.SET_READS, [mb["query"].tobytes()])
        d = engine_descs(G, mb)
        sync = [eng.align_tiles(d[k:k + 1000]) for k in range(0, 3000, 1000)]
        eng.submit(d[0:1000]); eng.submit(d[1000:2000])
        a = eng.wait(); eng.submit(d[2000:3000]); b = eng.wait(); c = eng.wait()
        from helpers import unpack_all
        for (r1, s1), (r2, s2) in zip(sync, [a, b, c]):
            assert (r1 == r2).all()
            assert (unpack_all(s1, r1["n_states"], 400) == unpack_all(s2, r2["n_states"], 400)).all()
        with pytest.raises(G.GactError):
            eng.wait()
        # the whole ring in flight (kernels of consecutive batches may overlap on the device), one more is refused
        assert G.MAX_INFLIGHT == 3
        for k in range(0, 3000, 1000):
            eng.submit(d[k:k + 1000])
        with pytest.raises(G.GactError):
            eng.submit(d[0:10])
        ring = [eng.wait() for _ in range(3)]
        for (r1, s1), (r2, s2) in zip(sync, ring):
            assert (r1 == r2).all()
            assert (unpack_all(s1, r1["n_states"], 400) == unpack_all(s2, r2["n_states"], 400)).all()
        eng.stage(d[:512]); eng.run_staged(); eng.run_staged()
        ms = eng.last_kernel_ms()
        r3, s3 = eng.fetch_staged()
        assert ms > 0 and (r3 == sync[0][0][:512]).all()
        st = eng.stats()
        assert st["tiles"] >= 3000 and st["cells"] > 0 and st["kernel_ms"] > 0


def test_descriptor_validation(pygact):
    G = pygact
    with G.GactEngine(max_tiles=8) as eng:
        eng.upload(G.SET_REF, [b"ACGT" * 10])
        eng.upload(G.SET_READS, [b"ACGT" * 10])
        d = G.make_descs(1)
        d["ref_len"], d["query_len"], d["query_set"] = 41, 4, G.SET_READS
        with pytest.raises(G.GactError):
            eng.align_tiles(d)
        with pytest.raises(G.GactError):
            eng.align_tiles(G.make_descs(9))


def test_large_batch_properties(pygact, oracle):
    """Full-size batch: exact on a sampled subset, plus size-independent invariants on all tiles."""
    G, O = pygact, oracle
    import synth
    n = 1 << 17
    mb = synth.tile_microbatch(n, seed=42)
    with G.GactEngine(max_tiles=n) as eng:
        eng.upload(G.SET_REF, [mb["ref"].tobytes()])
        eng.upload(G.SET_READS, [mb["query"].tobytes()])
        d = engine_descs(G, mb)
        res, st = eng.align_tiles(d)
        res2, st2 = eng.align_tiles(d)
    from helpers import unpack_all
    assert (res == res2).all()                                            # deterministic
    assert (unpack_all(st, res["n_states"], 400) == unpack_all(st2, res2["n_states"], 400)).all()
    assert (res["n_states"] <= 2 * 200 - 1).all()
    assert (np.maximum(res["i_steps"], res["j_steps"]) <= 200).all()
    assert (res["i_steps"] <= mb["ref_len"]).all() and (res["j_steps"] <= mb["query_len"]).all()
    nf = mb["first"] == 0
    assert (res["max_i"][nf] == mb["ref_len"][nf]).all() and (res["max_j"][nf] == mb["query_len"][nf]).all()
    u = unpack_all(st, res["n_states"], 400)
    assert ((u == 3) | (u == 2)).sum(axis=1).tolist() == res["i_steps"].tolist()
    assert ((u == 3) | (u == 1)).sum(axis=1).tolist() == res["j_steps"].tolist()
    sub = np.random.default_rng(0).choice(n, size=6000, replace=False)
    od = oracle_descs(O, mb)[sub]
    ores, ost = O.align_batch(mb["ref"], mb["query"], od, n_threads=8)
    assert len(compare_batch(res[sub], st[sub], ores, ost)) == 0
