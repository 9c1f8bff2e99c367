"""CPU-only: the oracle's GACT() restatement (oracle_gact_extend, gact.cpp:48-228) pinned to the reference CPU
build.  Candidates come from the host D-SOFT (pinned to the reference by tests/test_host.py), are converted as in
darwin.cpp:215-224 / 254-263, extended by the oracle, formatted as in gact.cpp:214-224, and the sorted|uniq lines
must equal the golden output of the unmodified reference (`darwin_ref`, tests/golden/e2e_*/expected_*.txt)."""
import os

import numpy as np
import pytest

from test_host import GOLD, host, read_fasta_simple  # noqa: F401  (host is a fixture)

CASES = [("e2e_acgt", "t320", (1, -1, -1, -1), 320, 120), ("e2e_acgt", "t512", (1, -1, -1, -1), 512, 192),
         ("e2e_acgt", "t200_s2", (2, -3, -5, -2), 200, 60), ("e2e_small", "t256", (1, -1, -1, -1), 256, 96),
         ("e2e_small", "t512_s2", (2, -3, -5, -2), 512, 192)]


@pytest.mark.parametrize("fixture,tag,scores,tile,overlap", CASES)
def test_oracle_gact_reproduces_reference_output(host, oracle, fixture, tag, scores, tile, overlap):  # noqa: F811
    import synth
    d = os.path.join(GOLD, fixture)
    refs = read_fasta_simple(os.path.join(d, "ref.fasta"))
    reads = read_fasta_simple(os.path.join(d, "reads.fasta"))
    bin_size = 64
    refstr, chr_start_bin, bin_to_chr = b"", [], []
    for i, (_, s) in enumerate(refs):                                  # darwin.cpp:530-543
        chr_start_bin.append(len(bin_to_chr))
        nb = (len(s) + bin_size - 1) // bin_size
        bin_to_chr += [i] * nb
        refstr += s + b"N" * (nb * bin_size - len(s))
    t = host.dh_seed_table_new(refstr, len(refstr), 14, 32, bin_size, 4, 2)
    assert t
    lines = set()
    for name, s in reads:
        rc = synth.revcomp(np.frombuffer(s, dtype=np.uint8)).tobytes()
        for strand, comp in ((s, 0), (rc, 1)):
            buf = np.zeros(4096, dtype=np.uint64)
            n = host.dh_dsoft(t, strand, len(strand), 800, 21, 1000000, 2500000, buf.ctypes.data, 4096)
            for c in buf[:n]:
                hit, qpos = int(c) >> 32, int(c) & 0xffffffff
                chrom = bin_to_chr[hit // bin_size]
                rpos = min(hit - chr_start_bin[chrom] * bin_size, len(refs[chrom][1]))      # darwin.cpp:222-224
                a, _ = oracle.gact_extend(refs[chrom][1], strand, rpos, qpos, tile_size=tile, tile_overlap=overlap,
                                          thr=35, scores=scores)
                if a.score > 0:                                                            # gact.cpp:213
                    lines.add(f"ref_id: {refs[chrom][0].split()[0]}, query_id: {name.split()[0]}, ab: {a.ab}, ae: {a.ae}, "
                              f"bb: {a.bb}, be: {a.be}, score: {a.score}, comp: {comp}")
    host.dh_seed_table_free(t)
    exp = open(os.path.join(d, f"expected_{tag}.txt")).read().splitlines()
    assert sorted(lines) == exp
