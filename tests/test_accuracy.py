"""tools/accuracy_report.py: the sweep over read positions equals the all-pairs definition of the reference's
accuracy script, and TP / FN / FP follow its rules (mirrored overlaps, self pairs dropped, length and score filter)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import accuracy_report as A  # noqa: E402


def brute(reads, min_overlap):
    out = set()
    for a in reads:
        for b in reads:
            if a[0] == b[0] or a[1] != b[1]:
                continue
            if min(a[2] + a[3], b[2] + b[3]) - max(a[2], b[2]) >= min_overlap:
                out.add((a[0], b[0]))
    return out


def test_theoretical_overlaps_sweep_equals_all_pairs():
    rng = np.random.default_rng(5)
    for n, span in ((1, 10000), (50, 20000), (400, 200000), (300, 30000)):
        reads = [(i, int(rng.integers(0, 3)), int(rng.integers(0, span)), int(rng.integers(200, 6000))) for i in range(n)]
        assert A.theoretical_overlaps(reads, 1000) == brute(reads, 1000)
    assert A.parse_name("S12_3400_5000") == (12, 0, 3400, 5000)
    assert A.parse_name("S7_2_100_900") == (7, 2, 100, 900)


def test_report_counts(tmp_path):
    fa = tmp_path / "reads.fasta"
    fa.write_text(">S0_0_0_5000\nACGT\n>S1_0_3000_5000\nACGT\n>S2_0_7500_3000\nACGT\n>S3_0_40000_2000\nACGT\n")
    # theoretical (>= 1000 shared): 0-1 (2000 shared); 1-2 shares 500 -> no; 0-2 none; 3 isolated
    ov = tmp_path / "out.darwin"
    ov.write_text(
        "ref_id: S0_0_0_5000, query_id: S1_0_3000_5000, ab: 3000, ae: 4990, bb: 5, be: 1995, score: 900, comp: 0\n"   # TP (x2 mirrored)
        "ref_id: S0_0_0_5000, query_id: S0_0_0_5000, ab: 0, ae: 5000, bb: 0, be: 5000, score: 5000, comp: 0\n"       # self pair dropped
        "ref_id: S2_0_7500_3000, query_id: S3_0_40000_2000, ab: 0, ae: 1500, bb: 0, be: 1500, score: 700, comp: 1\n" # FP (x2)
        "ref_id: S1_0_3000_5000, query_id: S2_0_7500_3000, ab: 0, ae: 1500, bb: 0, be: 1500, score: 599, comp: 0\n"  # score filter
        "ref_id: S1_0_3000_5000, query_id: S0_0_0_5000, ab: 0, ae: 900, bb: 0, be: 1500, score: 800, comp: 0\n")     # length filter
    r = A.report(str(fa), str(ov))
    assert r["theoretical_overlaps"] == 2 and r["overlap_lines"] == 5
    assert (r["TP"], r["FN"], r["FP"]) == (2, 0, 2)
    assert r["sensitivity"] == 1.0 and r["specificity"] == 0.5
    r = A.report(str(fa), str(ov), mirror=False)
    assert (r["TP"], r["FN"], r["FP"]) == (1, 1, 1)
