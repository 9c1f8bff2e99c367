"""Sensitivity / specificity of an overlap file against simulated read positions.

Python-3 counterpart of the reference's de-novo accuracy check (measure_sensitivity_PBSIM.py, README:26), same
definitions so the numbers are comparable:
  * reads are named  S<i>_<pos>_<len>  (the reference's convention) or  S<i>_<chr>_<pos>_<len>  (synth.py);
  * theoretical overlaps: ordered pairs (a, b), a != b, same sequence, whose genome intervals share >= 1000 bases;
  * reported overlaps: lines of `cat darwin.*.out | sort | uniq`, each also counted mirrored (query as reference),
    self pairs dropped, kept when both aligned spans are >= 990 bases and score >= 600;
  * TP = kept reported overlaps whose ordered read pair is a theoretical overlap, FP = the others,
    FN = theoretical overlaps without any kept reported overlap;
  * sensitivity = TP / (TP + FN), specificity = TP / (TP + FP)  (the reference's names for these two ratios).
Theoretical overlaps come from a sort + sweep over read start positions instead of the all-pairs loop.

    python tools/accuracy_report.py --reads reads.fasta --overlaps out.darwin
"""
import argparse
import json
import re
import sys

LINE = re.compile(r"ref_id: (\S+), query_id: (\S+), ab: (-?\d+), ae: (-?\d+), bb: (-?\d+), be: (-?\d+), score: (-?\d+), comp: (\d)")


def parse_name(name):
    """-> (read index, sequence index, start, length)"""
    v = [int(x) for x in re.findall(r"\d+", name)]
    if len(v) >= 4:
        return v[0], v[1], v[2], v[3]
    if len(v) == 3:
        return v[0], 0, v[1], v[2]
    raise ValueError(f"read name {name!r} carries no simulated position")


def read_names(fasta_path):
    names = []
    with open(fasta_path, "rb") as f:
        for ln in f:
            if ln.startswith(b">"):
                names.append(ln[1:].split()[0].decode())
    return names


def theoretical_overlaps(reads, min_overlap=1000):
    """reads: list of (idx, chr, start, length).  Ordered pairs (a, b), a != b, overlapping by >= min_overlap."""
    pairs = set()
    by_chr = {}
    for r in reads:
        by_chr.setdefault(r[1], []).append(r)
    for rs in by_chr.values():
        rs.sort(key=lambda r: r[2])
        active = []                                   # reads whose interval may still reach the current start
        for r in rs:
            a1, a2 = r[2], r[2] + r[3]
            active = [q for q in active if q[2] + q[3] - a1 >= min_overlap]
            for q in active:
                if min(a2, q[2] + q[3]) - max(a1, q[2]) >= min_overlap and q[0] != r[0]:
                    pairs.add((q[0], r[0]))
                    pairs.add((r[0], q[0]))
            active.append(r)
    return pairs


def reported_overlaps(path, mirror=True, score_thres=600, min_length=990):
    kept = []
    n_lines = 0
    with open(path) as f:
        for ln in f:
            m = LINE.search(ln)
            if not m:
                continue
            n_lines += 1
            ref_id, query_id = parse_name(m.group(1))[0], parse_name(m.group(2))[0]
            ab, ae, bb, be, score = (int(m.group(k)) for k in range(3, 8))
            cands = [(ref_id, query_id, ab, ae, bb, be, score)]
            if mirror:
                cands.append((query_id, ref_id, bb, be, ab, ae, score))
            for c in cands:
                if c[0] != c[1] and c[3] - c[2] >= min_length and c[5] - c[4] >= min_length and c[6] >= score_thres:
                    kept.append(c)
    return kept, n_lines


def report(reads_fasta, overlaps_path, mirror=True, score_thres=600, min_length=990, min_overlap=1000):
    reads = [parse_name(n) for n in read_names(reads_fasta)]
    tovl = theoretical_overlaps(reads, min_overlap)
    hovl, n_lines = reported_overlaps(overlaps_path, mirror, score_thres, min_length)
    hit = set()
    tp = fp = 0
    for h in hovl:
        if (h[0], h[1]) in tovl:
            tp += 1
            hit.add((h[0], h[1]))
        else:
            fp += 1
    fn = len(tovl) - len(hit)
    return {"reads": len(reads), "theoretical_overlaps": len(tovl), "overlap_lines": n_lines,
            "reported_after_filter": len(hovl), "TP": tp, "FN": fn, "FP": fp,
            "sensitivity": tp / (tp + fn) if tp + fn else None,
            "specificity": tp / (tp + fp) if tp + fp else None,
            "pairs_found": len(hit), "pair_recall": len(hit) / len(tovl) if tovl else None,
            "filter": {"score_thres": score_thres, "min_length": min_length, "min_overlap": min_overlap, "mirror": mirror}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", required=True)
    ap.add_argument("--overlaps", required=True, help="sorted|uniq concatenation of darwin.*.out")
    ap.add_argument("--score-thres", type=int, default=600)
    ap.add_argument("--min-length", type=int, default=990)
    ap.add_argument("--min-overlap", type=int, default=1000)
    ap.add_argument("--no-mirror", action="store_true")
    a = ap.parse_args()
    json.dump(report(a.reads, a.overlaps, not a.no_mirror, a.score_thres, a.min_length, a.min_overlap), sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
