"""CPU-only: host-side logic of the drop-in CLI (params.cfg, FASTA, minimizer seed table, D-SOFT)
against the reference's golden candidates.  No GACT compute happens here."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
HOST_SO = os.path.join(ROOT, "darwin-gpu_b200", "libdarwin_host.so")


@pytest.fixture(scope="module")
def host():
    if not os.path.exists(HOST_SO) or not os.path.exists(os.path.join(ROOT, "darwin-gpu_b200", "libgact_b200.so")):
        import __graft_entry__
        __graft_entry__.build()
    H = C.CDLL(HOST_SO)
    H.dh_seed_table_new.restype = C.c_void_p
    H.dh_seed_table_new.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    H.dh_seed_table_free.argtypes = [C.c_void_p]
    H.dh_dsoft.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    H.dh_hash32.restype = C.c_uint32
    H.dh_hash32.argtypes = [C.c_uint32, C.c_int]
    H.dh_read_fasta.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int]
    H.dh_params.argtypes = [C.c_char_p, C.c_void_p]
    return H


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


def test_params_cfg(host, tmp_path):
    p = tmp_path / "params.cfg"
    p.write_text("[GACT_scoring]\nmatch = 2\nmismatch= -3\ngap_open =-5\ngap_extend = -2\n# comment\n"
                 "[DSOFT_params]\nseed_size  = 14\nbin_size   = 64\nwindow_size= 4\nthreshold  = 21\nnum_seeds  = 800\n"
                 "seed_occurence_multiple = 32\nmax_candidates = 1000000\nnum_nz_bins    = 2500000\n\n"
                 "[GACT_first_tile]\nfirst_tile_size = 128\nfirst_tile_score_threshold = 35.7\n"
                 "[GACT_extend]\ntile_size = 512\ntile_overlap = 192\n")
    out = (C.c_int * 16)()
    assert host.dh_params(str(p).encode(), out) == 0
    assert list(out) == [2, -3, -5, -2, 14, 64, 4, 21, 800, 32, 1000000, 2500000, 128, 35, 512, 192]
    (tmp_path / "bad.cfg").write_text("[GACT_scoring]\nmatch = 1\n")
    assert host.dh_params(str(tmp_path / "bad.cfg").encode(), out) == -1        # missing keys are an error


def test_fasta_names_and_wrap_rule(host, tmp_path):
    names = C.create_string_buffer(4096)
    lens = (C.c_longlong * 64)()
    n = host.dh_read_fasta(os.path.join(GOLD, "e2e_small", "ref.fasta").encode(), names, 4096, lens, 64)
    assert n == 2 and names.value.decode().split("\n")[:2] == ["chrA", "chrB"] and list(lens[:2]) == [60000, 45000]
    bad = tmp_path / "bad.fasta"
    bad.write_text(">x\n" + "A" * 60 + "\n" + "C" * 60 + "\n")
    assert host.dh_read_fasta(str(bad).encode(), names, 4096, lens, 64) == -1   # short line after a short line
    bad.write_text(">x\n" + "A" * 71 + "\n")
    assert host.dh_read_fasta(str(bad).encode(), names, 4096, lens, 64) == -1   # longer than 70 columns
    ok = tmp_path / "ok.fasta"
    ok.write_text(">r1|extra\n" + "A" * 70 + "\n" + "C" * 12 + "\n>r_2 more\n" + "G" * 5 + "\n")
    n = host.dh_read_fasta(str(ok).encode(), names, 4096, lens, 64)
    assert n == 2 and names.value.decode().split("\n")[:2] == ["r1", "r_2"] and list(lens[:2]) == [82, 5]


def test_hash32_known_values(host):
    # values of the reference's hash32(key, 14) (ntcoding.cpp:74-85), computed from its published formula
    def h(key, k):
        m = (1 << 2 * k) - 1
        M32 = 0xFFFFFFFF
        key = ((~key & M32) + ((key << 21) & M32)) & M32 & m
        key ^= key >> 24
        key = ((key + ((key << 3) & M32) & M32) + ((key << 8) & M32)) & M32 & m
        key ^= key >> 14
        key = (((key + ((key << 2) & M32)) & M32) + ((key << 4) & M32)) & M32 & m
        key ^= key >> 28
        key = (key + ((key << 31) & M32)) & M32 & m
        return key
    for key in [0, 1, 12345, (1 << 28) - 1, 0x0ABCDEF]:
        assert host.dh_hash32(key, 14) == h(key, 14)


def test_dsoft_matches_reference_golden(host):
    """Candidate stream for every read/strand of e2e_small equals the reference's DSOFT."""
    import synth
    z = np.load(os.path.join(GOLD, "dsoft_e2e_small.npz"))
    refs = read_fasta_simple(os.path.join(GOLD, "e2e_small", "ref.fasta"))
    reads = read_fasta_simple(os.path.join(GOLD, "e2e_small", "reads.fasta"))
    refstr = b"".join(s + b"N" * ((64 - len(s) % 64) % 64) for _, s in refs)
    t = host.dh_seed_table_new(refstr, len(refstr), 14, 32, 64, 4, 2)
    assert t
    got_counts, got = [], []
    for _, s in reads:
        for strand in (s, synth.revcomp(np.frombuffer(s, dtype=np.uint8)).tobytes()):
            buf = np.zeros(4096, dtype=np.uint64)
            n = host.dh_dsoft(t, strand, len(strand), 800, 21, 1000000, 2500000, buf.ctypes.data, 4096)
            got_counts.append(n)
            got.append(buf[:n].copy())
    host.dh_seed_table_free(t)
    assert got_counts == z["counts"].tolist()
    assert (np.concatenate(got) == z["cands"]).all()


def test_seed_table_matches_reference_golden(host):
    """index_table_ / pos_table_ of the host builder equal the reference's own SeedPosTable constructor
    (seed_pos_table.cpp:46-98): SHA-256 digests committed by tests/golden/make_golden.py, and -- where the reference
    was built in place (oracle/_ref) -- the live tables word for word."""
    import json
    from helpers import seedtable_cases, table_digest
    host.dh_seed_table_arrays.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    gold = json.load(open(os.path.join(GOLD, "seedtable_digests.json")))
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libseed_ref.so")
    R = None
    if os.path.exists(ref_so):
        R = C.CDLL(ref_so)
        if hasattr(R, "ref_seed_table_arrays"):
            R.ref_seed_table_new.restype = C.c_void_p
            R.ref_seed_table_new.argtypes = [C.c_char_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
            R.ref_seed_table_arrays.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        else:
            R = None
    cases = seedtable_cases()
    assert sorted(gold) == sorted(c[0] for c in cases)
    for tag, refstr, k, w, b in cases:
        t = host.dh_seed_table_new(refstr, len(refstr), k, 32, b, w, 4)
        assert t
        ip, ie, pp, npos, mo = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64(), C.c_uint32()
        host.dh_seed_table_arrays(t, C.byref(ip), C.byref(ie), C.byref(pp), C.byref(npos), C.byref(mo))
        index = np.ctypeslib.as_array((C.c_uint32 * ie.value).from_address(ip.value))
        pos = np.ctypeslib.as_array((C.c_uint32 * max(npos.value, 1)).from_address(pp.value))[:npos.value]
        d = table_digest(index, pos)
        for key in ("index_entries", "n_pos", "index_sha256", "pos_sha256"):
            assert d[key] == gold[tag][key], (tag, key)
        if R is not None:
            rt = R.ref_seed_table_new(refstr, len(refstr), k, 32, b, w)
            rip, rie, rpp, rn = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint64()
            R.ref_seed_table_arrays(rt, C.byref(rip), C.byref(rie), C.byref(rpp), C.byref(rn))
            assert rie.value == ie.value and rn.value == npos.value
            assert np.array_equal(np.ctypeslib.as_array((C.c_uint32 * rie.value).from_address(rip.value)), index)
            assert np.array_equal(np.ctypeslib.as_array((C.c_uint32 * max(rn.value, 1)).from_address(rpp.value))[:rn.value], pos)
        host.dh_seed_table_free(t)


def test_cli_usage_and_no_gpu_behaviour(tmp_path):
    """The drop-in binary keeps the reference's usage line; without a GPU it refuses to run
    (no CPU fallback on the GACT path)."""
    exe = os.path.join(ROOT, "darwin-gpu_b200", "darwin")
    if not os.path.exists(exe):
        import __graft_entry__
        __graft_entry__.build()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage: ./darwin <REFERENCE>.fasta <READS>.fasta CPU_THREADS" in r.stderr
    import pygact
    if pygact.device_count() > 0:
        pytest.skip("GPU present")
    (tmp_path / "params.cfg").write_text(open(os.path.join(ROOT, "darwin-gpu_b200", "params.cfg")).read())
    r = subprocess.run([exe, os.path.join(GOLD, "e2e_small", "ref.fasta"), os.path.join(GOLD, "e2e_small", "reads.fasta"), "2"],
                       capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 2 and "no CUDA device" in r.stderr
