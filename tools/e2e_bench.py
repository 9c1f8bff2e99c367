"""End-to-end reads/s measurement of the drop-in `darwin` binary (BASELINE.json configs 3/4/5 shape).

Generates a seeded synthetic reference + PacBio-like reads (synth.py), runs darwin-gpu_b200/darwin on all of
it, optionally runs the reference CPU build (oracle/_ref/darwin_ref) on the first --ref-reads reads with all
host cores, and checks that the sorted|uniq outputs of that subset are byte-identical.

    python tools/e2e_bench.py --ref-mbp 100 --reads-mb 50 --gpus 1 --ref-reads 96 --out gpurun_out/e2e.json
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "darwin-gpu_b200"))
import numpy as np  # noqa: E402
import synth  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tools"))
import accuracy_report  # noqa: E402

PARAMS = open(os.path.join(ROOT, "darwin-gpu_b200", "params.cfg")).read()


def collect(workdir):
    lines = []
    for fn in sorted(os.listdir(workdir)):
        if fn.startswith("darwin.") and fn.endswith(".out"):
            lines += open(os.path.join(workdir, fn)).read().splitlines()
            os.remove(os.path.join(workdir, fn))
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-mbp", type=float, default=20.0)
    ap.add_argument("--ref-pieces", type=int, default=20)
    ap.add_argument("--reads-mb", type=float, default=10.0)
    ap.add_argument("--read-mean", type=float, default=10000.0)
    ap.add_argument("--read-sd", type=float, default=3000.0)
    ap.add_argument("--read-max", type=int, default=30000)
    ap.add_argument("--tile", type=int, default=320)
    ap.add_argument("--overlap", type=int, default=120)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--ref-reads", type=int, default=0, help="reads for the reference CPU build arm (0 = skip)")
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--self-align", action="store_true",
                    help="config 1 shape: de-novo self alignment (reads.fasta against itself, PBSIM-CLR-like 3 kb reads)")
    ap.add_argument("--out", default="")
    ap.add_argument("--show-stdout", action="store_true", help="print the timing lines of the darwin run")
    args = ap.parse_args()

    wd = tempfile.mkdtemp(prefix="darwin_e2e_")
    rng = np.random.default_rng(args.seed)
    piece = int(args.ref_mbp * 1e6 / args.ref_pieces)
    genome = [synth.random_genome(piece, rng) for _ in range(args.ref_pieces)]
    synth.write_fasta(os.path.join(wd, "ref.fasta"), [f"chr{i}" for i in range(args.ref_pieces)], genome)
    names, reads = synth.sample_reads(genome, int(args.reads_mb * 1e6), np.random.default_rng(args.seed + 1),
                                      mean=args.read_mean, sd=args.read_sd, lo=1000, hi=args.read_max)
    synth.write_fasta(os.path.join(wd, "reads.fasta"), names, reads)
    cfg = re.sub(r"tile_size = \d+", f"tile_size = {args.tile}", PARAMS)
    cfg = re.sub(r"tile_overlap = \d+", f"tile_overlap = {args.overlap}", cfg)
    open(os.path.join(wd, "params.cfg"), "w").write(cfg)
    n_reads, n_bases = len(reads), int(sum(len(r) for r in reads))

    ref_name = "reads.fasta" if args.self_align else "ref.fasta"      # same file name on both sides -> same_file (darwin.cpp:498-503)
    env = dict(os.environ, DARWIN_GPUS=str(args.gpus))
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(ROOT, "darwin-gpu_b200", "darwin"), ref_name, "reads.fasta", str(args.threads)],
                       cwd=wd, capture_output=True, text=True, env=env)
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:])
        raise SystemExit("darwin failed")
    ours = collect(wd)
    if args.show_stdout:
        print("\n".join(ln for ln in r.stdout.splitlines() if re.search(r"Time |init|build|driver|num_candidates|Shard", ln)), f"\nwall {wall:.3f} s")
    summ = json.loads(re.search(r"DARWIN_B200_SUMMARY (\{.*\})", r.stdout).group(1))
    phase = {k: int(v) for k, v in re.findall(r"Time elapsed \(([^)]*)\): (\d+) msec", r.stdout)}
    seeds_ms = [int(x) for x in re.findall(r"Time finding seeds: (\d+) msec", r.stdout)]
    init_ms = [int(x) for x in re.findall(r"init alone[^:]*: (\d+) msec", r.stdout)]
    table_ms = [float(x) for x in re.findall(r"seed table build[^:]*: ([\d.]+) msec", r.stdout)]
    align_s = summ["align_phase_ms"] / 1e3
    res = {"workload": {"ref_mbp": args.ref_mbp, "reads": n_reads, "read_bases": n_bases, "tile_size": args.tile,
                        "tile_overlap": args.overlap, "seed": args.seed},
           "gpus": args.gpus, "host_threads": args.threads,
           "ours": {"wall_s": wall, "phases_ms": phase, "dsoft_ms_per_shard": seeds_ms, "gpu_init_alone_ms": init_ms, "gpu_seed_table_build_ms": table_ms, "summary": summ,
                    "reads_per_s_align_phase": n_reads / align_s,
                    "reads_per_s_gact_only": n_reads / max(summ["gact_sched_ms"] / 1e3, 1e-9),
                    "gcups_align_phase": summ["cells"] / align_s / 1e9,
                    "gcups_gact_sched": summ["cells"] / max(summ["gact_sched_ms"] / 1e3, 1e-9) / 1e9,
                    "gcups_kernel": summ["cells"] / max(summ["gact_kernel_ms"] / 1e3, 1e-9) / 1e9,
                    "overlap_lines": len(ours), "unique_lines": len(set(ours))}}

    if args.self_align:
        # accuracy of the de-novo overlaps against the simulated read positions (reference: measure_sensitivity_PBSIM.py)
        with open(os.path.join(wd, "out.darwin"), "w") as f:
            f.write("".join(ln + "\n" for ln in sorted(set(ours))))
        res["ours"]["accuracy"] = accuracy_report.report(os.path.join(wd, "reads.fasta"), os.path.join(wd, "out.darwin"))

    ref_exe = os.path.join(ROOT, "oracle", "_ref", "darwin_ref")
    if args.ref_reads > 0 and os.path.exists(ref_exe):
        k = n_reads if args.self_align else min(args.ref_reads, n_reads)
        synth.write_fasta(os.path.join(wd, "reads_sub.fasta"), names[:k], reads[:k])
        t0 = time.perf_counter()
        ref_args = ["reads.fasta", "reads.fasta"] if args.self_align else ["ref.fasta", "reads_sub.fasta"]
        rr = subprocess.run([ref_exe, *ref_args, str(args.threads)], cwd=wd, capture_output=True, text=True)
        ref_wall = time.perf_counter() - t0
        ref_lines = collect(wd)
        ref_align = int(re.search(r"Time elapsed \(seed table querying \+ aligning\): (\d+) msec", rr.stdout).group(1)) / 1e3
        sub_names = set(names[:k])
        ours_sub = sorted({ln for ln in ours if re.search(r"query_id: (\S+),", ln).group(1) in sub_names})
        identical = ours_sub == sorted(set(ref_lines))
        res["reference_cpu"] = {"reads": k, "threads": args.threads, "wall_s": ref_wall, "align_phase_s": ref_align,
                                "reads_per_s_align_phase": k / ref_align, "unique_lines": len(set(ref_lines)),
                                "sorted_uniq_identical_to_ours_on_subset": identical}
        res["speedup_reads_per_s_align_phase"] = res["ours"]["reads_per_s_align_phase"] / (k / ref_align)
        if not identical:
            print("MISMATCH vs reference CPU build on the subset", file=sys.stderr)
    print(json.dumps(res, indent=1))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
