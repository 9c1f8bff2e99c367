"""Regenerates the golden fixtures from the UNMODIFIED reference built in place
(oracle/Makefile -> oracle/_ref/libalign_ref.so and oracle/_ref/darwin_ref).
Runs only in the build container (needs /root/reference); the fixtures it writes are committed.

  align_random.npz  -- random tiles through the reference's AlignWithBT (align.cpp:60-233)
  e2e_small/        -- a small read set + the reference CPU build's sorted|uniq output
                       (README:32 / x_scalingrun.sh:27-33 recipe), per tile_size config
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "darwin-gpu_b200")]
import oracle as O      # noqa: E402
import synth            # noqa: E402


def make_align_random():
    rng = np.random.default_rng(20261018)
    schemes = [(1, -1, -1, -1), (2, -3, -5, -2), (1, -1, -2, -1), (5, -4, -10, -1), (1, -3, 0, 0), (3, -1, -1, -4)]
    refs, queries, meta, queues = [], [], [], []
    for it in range(360):
        T = int(rng.choice([320, 320, 320, 64, 17, 1, 5, 200, 256]))
        sc = schemes[it % len(schemes)]
        full = it % 3 == 0
        rl = T if full else int(rng.integers(1, T + 1))
        ql = T if full else int(rng.integers(1, T + 1))
        g = synth.random_genome(rl, rng)
        q, _ = synth.error_channel(g, rng)
        q = np.concatenate([q, synth.random_genome(T, rng)])[:ql]
        if it % 11 == 0:
            q = np.frombuffer(b"ACGTNacgt", dtype=np.uint8)[rng.integers(0, 9, size=ql)]
        if it % 13 == 0:
            q = synth.random_genome(ql, rng)          # unrelated: low scores, ZERO states
        rev, first = it % 2, (it // 2) % 2
        et = int(rng.choice([200, 200, 4, 1, 50, 400]))
        ref_q = O.ref_align_tile(g.tobytes(), q.tobytes(), sc, rev, first, et)
        refs.append(g.tobytes()); queries.append(q.tobytes())
        meta.append(list(sc) + [rev, first, et])
        queues.append(np.asarray(ref_q, dtype=np.int32))
    np.savez_compressed(os.path.join(HERE, "align_random.npz"),
                        refs=np.array(refs, dtype=object), queries=np.array(queries, dtype=object),
                        meta=np.asarray(meta, dtype=np.int32), queues=np.array(queues, dtype=object),
                        allow_pickle=True)
    print("align_random.npz:", len(refs), "tiles")


PARAMS = """[GACT_scoring]
match = {ma}
mismatch = {mi}
gap_open = {go}
gap_extend = {ge}
[DSOFT_params]
seed_size = 14
bin_size = 64
window_size = 4
threshold = 21
num_seeds = 800
seed_occurence_multiple = 32
max_candidates = 1000000
num_nz_bins = 2500000
[GACT_first_tile]
first_tile_size = 128
first_tile_score_threshold = 35
[GACT_extend]
tile_size = {ts}
tile_overlap = {to}
"""


def run_ref(workdir, ref_fa, reads_fa, threads, cfg):
    os.makedirs(workdir, exist_ok=True)
    with open(os.path.join(workdir, "params.cfg"), "w") as f:
        f.write(PARAMS.format(**cfg))
    for fn in os.listdir(workdir):
        if fn.startswith("darwin.") and fn.endswith(".out"):
            os.remove(os.path.join(workdir, fn))
    subprocess.run([O.REF_DARWIN, ref_fa, reads_fa, str(threads)], cwd=workdir, check=True,
                   stdout=subprocess.DEVNULL)
    lines = []
    for fn in sorted(os.listdir(workdir)):
        if fn.startswith("darwin.") and fn.endswith(".out"):
            lines += open(os.path.join(workdir, fn)).read().splitlines()
    return sorted(set(lines))


def make_e2e_small():
    out = os.path.join(HERE, "e2e_small")
    os.makedirs(out, exist_ok=True)
    rng = np.random.default_rng(7)
    genome = [synth.random_genome(60000, rng), synth.random_genome(45000, rng)]
    synth.write_fasta(os.path.join(out, "ref.fasta"), ["chrA extra words", "chrB"], genome)
    names, reads = synth.sample_reads(genome, 160000, rng, mean=3000, sd=1500, lo=300, hi=9000)
    # a lower-case / N-containing read exercises the raw-byte comparison (align.cpp:134)
    r = reads[3].copy(); r[100:140] = np.frombuffer(b"N", dtype=np.uint8)[0]; r[500:520] += 32; reads[3] = r
    synth.write_fasta(os.path.join(out, "reads.fasta"), names, reads)
    tmp = "/tmp/golden_e2e"
    cfgs = {"t320": dict(ma=1, mi=-1, go=-1, ge=-1, ts=320, to=120),
            "t256": dict(ma=1, mi=-1, go=-1, ge=-1, ts=256, to=96),
            "t512_s2": dict(ma=2, mi=-3, go=-5, ge=-2, ts=512, to=192)}
    for tag, cfg in cfgs.items():
        lines = run_ref(tmp, os.path.join(out, "ref.fasta"), os.path.join(out, "reads.fasta"), 3, cfg)
        open(os.path.join(out, f"expected_{tag}.txt"), "w").write("\n".join(lines) + "\n")
        print(tag, "reads-vs-ref lines:", len(lines))
    # de-novo self alignment (same file name on both sides -> same_file suppression, gact.cpp:213)
    lines = run_ref(tmp, os.path.join(out, "reads.fasta"), os.path.join(out, "reads.fasta"), 3, cfgs["t320"])
    open(os.path.join(out, "expected_self_t320.txt"), "w").write("\n".join(lines) + "\n")
    print("self lines:", len(lines))


def make_e2e_acgt():
    """ACGT-only reads (2-bit packed sets on the device: the on-device chain + D-SOFT path of the CLI)."""
    out = os.path.join(HERE, "e2e_acgt")
    os.makedirs(out, exist_ok=True)
    rng = np.random.default_rng(11)
    genome = [synth.random_genome(70000, rng), synth.random_genome(30000, rng)]
    synth.write_fasta(os.path.join(out, "ref.fasta"), ["g1", "g2_b"], genome)
    names, reads = synth.sample_reads(genome, 150000, rng, mean=4000, sd=2500, lo=200, hi=12000)
    synth.write_fasta(os.path.join(out, "reads.fasta"), names, reads)
    tmp = "/tmp/golden_e2e_acgt"
    cfgs = {"t320": dict(ma=1, mi=-1, go=-1, ge=-1, ts=320, to=120),
            "t512": dict(ma=1, mi=-1, go=-1, ge=-1, ts=512, to=192),
            "t1024": dict(ma=1, mi=-1, go=-1, ge=-1, ts=1024, to=384),
            "t200_s2": dict(ma=2, mi=-3, go=-5, ge=-2, ts=200, to=60)}
    for tag, cfg in cfgs.items():
        lines = run_ref(tmp, os.path.join(out, "ref.fasta"), os.path.join(out, "reads.fasta"), 4, cfg)
        open(os.path.join(out, f"expected_{tag}.txt"), "w").write("\n".join(lines) + "\n")
        print("acgt", tag, "lines:", len(lines))


def make_dsoft_golden():
    """Candidates of the reference's own SeedPosTable::DSOFT (seed_pos_table.cpp:100-167) for the
    e2e_small reads (both strands) against the e2e_small reference."""
    import ctypes as C
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libseed_ref.so"))
    R.ref_seed_table_new.restype = C.c_void_p
    R.ref_seed_table_new.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
    R.ref_dsoft.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int,
                            C.c_int, C.c_void_p, C.c_int]
    out = os.path.join(HERE, "e2e_small")
    refs = read_fasta_simple(os.path.join(out, "ref.fasta"))
    reads = read_fasta_simple(os.path.join(out, "reads.fasta"))
    refstr = b""
    for _, s in refs:
        refstr += s + (b"N" * ((64 - len(s) % 64) % 64))
    t = R.ref_seed_table_new(refstr, len(refstr), 14, 32, 64, 4)
    counts, cands = [], []
    for _, s in reads:
        for strand in (s, synth.revcomp(np.frombuffer(s, dtype=np.uint8)).tobytes()):
            buf = np.zeros(4096, dtype=np.uint64)
            n = R.ref_dsoft(t, strand, len(strand), len(refstr), 64, 800, 21, 1000000, 2500000, buf.ctypes.data, 4096)
            counts.append(n)
            cands.append(buf[:n].copy())
    np.savez_compressed(os.path.join(HERE, "dsoft_e2e_small.npz"), counts=np.asarray(counts, dtype=np.int32),
                        cands=np.concatenate(cands) if cands else np.zeros(0, dtype=np.uint64))
    print("dsoft golden:", int(sum(counts)), "candidates over", len(counts), "strand calls")


def make_seedtable_golden():
    """SHA-256 of index_table_ / pos_table_ as the reference's own SeedPosTable constructor builds them
    (seed_pos_table.cpp:46-98) for the inputs of tests/helpers.py:seedtable_cases()."""
    import ctypes as C
    import json
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libseed_ref.so"))
    R.ref_seed_table_new.restype = C.c_void_p
    R.ref_seed_table_new.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
    R.ref_seed_table_arrays.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    out = {}
    for tag, refstr, k, w, b in helpers.seedtable_cases():
        t = R.ref_seed_table_new(refstr, len(refstr), k, 32, b, w)
        ip, ie, pp, npos = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64()
        R.ref_seed_table_arrays(t, C.byref(ip), C.byref(ie), C.byref(pp), C.byref(npos))
        index = np.ctypeslib.as_array((C.c_uint32 * ie.value).from_address(ip.value))
        pos = np.ctypeslib.as_array((C.c_uint32 * max(npos.value, 1)).from_address(pp.value))[:npos.value]
        out[tag] = helpers.table_digest(index, pos)
        out[tag].update({"k": k, "w": w, "bin_size": b, "ref_len": len(refstr)})
        print("seed table golden", tag, out[tag]["n_pos"], "minimizers")
    json.dump(out, open(os.path.join(HERE, "seedtable_digests.json"), "w"), indent=1)


def make_tiny_golden():
    """Queues of the reference's own AlignWithBT (align.cpp:60-233) for every 7th tile of the exhaustive tiny-tile
    batch (tests/helpers.py:tiny_tile_batch), three scoring schemes, early_terminate 2 and 8."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    mb = helpers.tiny_tile_batch()
    sub = helpers.tiny_golden_subset(len(mb["ref_off"]))
    flat, offs = [], [0]
    for scores in helpers.TINY_SCHEMES:
        for T, ov in helpers.TINY_ENGINES:
            for t in sub:
                r = mb["ref"][mb["ref_off"][t]:mb["ref_off"][t] + mb["ref_len"][t]].tobytes()
                q = mb["query"][mb["query_off"][t]:mb["query_off"][t] + mb["query_len"][t]].tobytes()
                queue = O.ref_align_tile(r, q, scores, int(mb["reverse"][t]), int(mb["first"][t]), T - ov)
                flat += queue
                offs.append(len(flat))
    np.savez_compressed(os.path.join(HERE, "tiny_tiles.npz"), flat=np.asarray(flat, dtype=np.int16),
                        offs=np.asarray(offs, dtype=np.int32))
    print("tiny golden:", len(offs) - 1, "reference queues")


def read_fasta_simple(path):
    recs, name, cur = [], None, []
    for ln in open(path, "rb").read().split(b"\n"):
        if ln.startswith(b">"):
            if name is not None:
                recs.append((name, b"".join(cur)))
            name, cur = ln[1:].decode(), []
        elif ln:
            cur.append(ln)
    if name is not None:
        recs.append((name, b"".join(cur)))
    return recs


if __name__ == "__main__":
    O.build(ref=True)
    make_align_random()
    make_e2e_small()
    make_e2e_acgt()
    make_dsoft_golden()
    make_seedtable_golden()
    make_tiny_golden()
