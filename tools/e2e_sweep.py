"""Config 3 / 4 workload once, then the `darwin` binary under several settings: prints the phase lines, the summary and the
timeline of every run (where the align phase and the whole-program wall time go).

  python tools/e2e_sweep.py <gpus,comma separated> [reads_mb] [env=VAL ...  (one run per '/'-separated group)]"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "darwin-gpu_b200"))
import numpy as np
import synth

gpus = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1").split(",")]
reads_mb = float(sys.argv[2]) if len(sys.argv) > 2 else 50.0
groups = [dict(kv.split("=", 1) for kv in g.split(",") if kv) for g in (sys.argv[3].split("/") if len(sys.argv) > 3 else [""])]
wd = tempfile.mkdtemp(prefix="e2e_sweep_")
rng = np.random.default_rng(3)
genome = [synth.random_genome(5_000_000, rng) for _ in range(20)]
synth.write_fasta(os.path.join(wd, "ref.fasta"), [f"chr{i}" for i in range(20)], genome)
names, reads = synth.sample_reads(genome, int(reads_mb * 1e6), np.random.default_rng(4), mean=10000, sd=3000, lo=1000, hi=30000)
synth.write_fasta(os.path.join(wd, "reads.fasta"), names, reads)
open(os.path.join(wd, "params.cfg"), "w").write(open(os.path.join(ROOT, "darwin-gpu_b200", "params.cfg")).read())
exe = os.path.join(ROOT, "darwin-gpu_b200", "darwin")
threads = len(os.sched_getaffinity(0))
base = None
for g in gpus:
    for env in groups:
        for rep in range(2):
            for fn in os.listdir(wd):
                if fn.startswith("darwin.") and fn.endswith(".out"):
                    os.remove(os.path.join(wd, fn))
            t0 = time.perf_counter()
            r = subprocess.run([exe, "ref.fasta", "reads.fasta", str(threads)], cwd=wd, capture_output=True, text=True,
                               env=dict(os.environ, DARWIN_GPUS=str(g), **env))
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                print("FAILED", g, env, r.stderr[-300:])
                continue
            lines = []
            for fn in sorted(os.listdir(wd)):
                if fn.startswith("darwin.") and fn.endswith(".out"):
                    lines += open(os.path.join(wd, fn)).read().splitlines()
            uniq = sorted(set(lines))
            if base is None:
                base = uniq
            summ = json.loads(re.search(r"DARWIN_B200_SUMMARY (\{.*\})", r.stdout).group(1))
            tl = json.loads(re.search(r"DARWIN_B200_TIMELINE (\{.*\})", r.stdout).group(1))
            print(f"gpus {g} env {env} rep {rep}: align_phase_ms {summ['align_phase_ms']:.2f} reads/s {len(reads) / summ['align_phase_ms'] * 1e3:.0f} "
                  f"wall_s {wall:.2f} same_output {uniq == base} lines {len(uniq)}")
            if rep == 1:
                for ln in r.stdout.splitlines():
                    if ln.startswith("Time ") or ln.startswith("GPU ") or ln.startswith("TRACE"):
                        print("    " + ln)
                print("    timeline", tl)
