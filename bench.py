#!/usr/bin/env python
"""bench.py -- GACT tile-alignment throughput on B200 (BASELINE.json metric, config 2).

A "step" is one pass of the GACT tile path over one batch of synthetic tiles: the config-2
microbatch (1 Mi independent tile_size=320 tile pairs, 15 % PacBio-like error, 82 % full /
18 % edge tiles, 5.5 % first tiles, seed 42) per GPU.  Weak scaling: every rank aligns its own
batch (reads partition by index, no collective on the data path; darwin.cpp:619-629).

  value  : whole-job GCUPS (sum ref_len*query_len over all ranks' tiles / max-over-ranks device
           time), descriptors and sequences already resident in HBM.
  e2e    : the same metric through the public C ABI with HOST buffers (gact_engine_submit /
           gact_engine_wait, chunked and double-buffered): descriptor H2D and result + state D2H
           are inside the timed region.
  roofline : achieved = cells/s x 2.5 ALU-pipe lane-ops per cell (algorithmic floor of the packed recurrence) against
           the ALU-pipe issue rate measured in this run; `executed` turns this round's ncu counters of the same kernel
           (profiles/r2_tile_kernel_ncu.json) and the measured tile rate into issue-slot / ALU-pipe utilisation.
  cpu_baseline / cpu_baseline_port / gpu_baseline (N = 1): the reference's own AlignWithBT, the allocation-free oracle
           port of it, and the reference's own GPU kernel compiled for sm_100a, all on the same box in the same run.
  reads    : the application metric -- reads/s of the drop-in `darwin` binary on the full config 3 / 4 workload at
           --gpus GPUs (strong scaling), a weak-scaled run for N > 1, and the reference CPU build on a 200-read subset
           with its sorted|uniq output compared with ours.

`--impl reference` times the reference's own CPU AlignWithBT (oracle/_ref, built in place from
/root/reference) -- or the oracle port when that library is absent -- on a bounded sample of the
same tiles with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(ROOT, "darwin-gpu_b200")]

import numpy as np  # noqa: E402

METRIC = "gact_gcups"
UNIT = "GCUPS"
OPS_PER_CELL_SURVEY = 17.0   # scalar int32 instructions per DP cell (kept for continuity with round 1)
# Algorithmic floor of the packed recurrence: per s16x2 cell PAIR one substitution select (PRMT), M = max(diag + s, 0),
# I = max(I + ge, Mup + go), D = max(D + ge, Mleft + go), H = max3(M, I, D) -> 5 ALU-pipe instructions (the two gap-open
# adds run as IMAD on the FMA pipe), i.e. 2.5 ALU-pipe lane-operations per cell; direction codes, traceback, staging and
# wavefront skew come on top, so achieved / peak computed from it cannot exceed 1.
ALU_OPS_PER_CELL_FLOOR = 2.5
# this round's ncu captures of the two kernels of a tile step: the inter-task kernel (full, non-first tiles) and the
# wavefront kernel (everything else)
NCU_SUMMARIES = [("inter_task", os.path.join(ROOT, "profiles", "r2_inter_task_kernel_ncu.json")),
                 ("wavefront", os.path.join(ROOT, "profiles", "r2_wavefront_kernel_ncu.json")),
                 ("wavefront_narrow", os.path.join(ROOT, "profiles", "r2_narrow_kernel_ncu.json"))]
TILE, OVERLAP = 320, 120
SCORES = (1, -1, -1, -1)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tiles", type=int, default=1 << 20, help="tiles per GPU per step")
    ap.add_argument("--chunk", type=int, default=1 << 18, help="tiles per submit() in the e2e leg")
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 int32, 2 s16x2")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reads-leg", action="store_true", help="skip the application-level reads/s leg")
    ap.add_argument("--config", type=int, default=0, choices=[0, 1, 5],
                    help="run one of the other BASELINE.json configs at size instead of the bench line: 1 = de-novo self-alignment "
                         "of a 10x read set of a 4.64 Mbp genome, 5 = tile-size sweep 256/512/1024 on 30 kb reads")
    ap.add_argument("--cpu-reads", type=int, default=100, help="--config: reads given to the reference CPU build")
    return ap.parse_args()


def config(args, extra=None):
    c = {"workload": "config2: tile microbatch, 1Mi tile pairs tile_size=320 overlap=120, 15% PacBio-like error, "
                     "82% full/18% edge, 5.5% first, seed 42 (+rank)",
         "tiles_per_gpu_per_step": args.tiles, "tile_size": TILE, "tile_overlap": OVERLAP,
         "scores": list(SCORES), "parallelism": f"reads/tiles sharded over {args.gpus} GPU(s), no collective",
         "l2_policy": "no explicit flush: at the default size one step touches more than the 126 MB L2 (descriptors + "
                      "tile order 36 MiB, packed bases of 1 Mi tile windows, results + states 128 MiB, direction-window "
                      "scratch 115 MB), and the kernel is ALU-bound at 0.003 algorithmic bytes per cell"}
    assert not extra, "both arms must print key-identical config dicts"
    return c


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []          # (arrival time, line)
        self.t_begin = None      # start of the timed region: earlier samples (warm-up) are not used

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def begin(self):
        """The timed region starts now.  nvidia-smi was started before the warm-up (it needs a few hundred ms to come up,
        as long as a short timed region lasts), so samples exist from the first 100 ms of the region on."""
        self.t_begin = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_end = time.time()
        inside = [ln for t, ln in self.lines if self.t_begin is None or self.t_begin <= t <= t_end]
        if not inside and self.lines:
            # a timed region shorter than one sampling period: the sample closest to it (the GPU runs the same kernels in the
            # warm-up just before)
            inside = [min(self.lines, key=lambda x: abs(x[0] - (self.t_begin or t_end)))[1]]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = [x for x in sm if x >= 0.5 * max(sm)]
        return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


# ---------------------------------------------------------------------------------------------
def usable_cores():
    """Host cores this process may really use: affinity mask and cgroup CPU quota, not just the machine's count
    (a thread pool wider than the quota gets throttled in bursts, which shows up as tens of ms of jitter)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // per))
        except Exception:
            pass
    return n


def make_batch(n_tiles, seed):
    import synth
    return synth.tile_microbatch(n_tiles, tile_size=TILE, seed=seed)


def cpu_arm(args, mb, seconds, gpu_scores=None, force_port=False):
    """Reference CPU AlignWithBT (oracle/_ref) or the allocation-free oracle port, all host threads, bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    cores = usable_cores()
    n = len(mb["ref_off"])
    od = np.zeros(n, dtype=O.TILE_DESC_DTYPE)
    for k in ("ref_off", "query_off", "ref_len", "query_len", "reverse", "first"):
        od[k] = mb[k]
    use_ref = O.ref_available() and not force_port

    scores = np.zeros(n, dtype=np.int32)

    def run(k):
        t0 = time.perf_counter()
        if use_ref:
            cells = O.ref_lib().ref_align_batch(mb["ref"].ctypes.data, mb["query"].ctypes.data, od[:k].ctypes.data, k,
                                                *SCORES, TILE - OVERLAP, cores, scores.ctypes.data)
        else:
            res, _ = O.align_batch(mb["ref"], mb["query"], od[:k], scores=SCORES, et=TILE - OVERLAP, max_len=TILE,
                                   n_threads=cores)
            scores[:k] = res["score"]
            cells = int((od["ref_len"][:k].astype(np.int64) * od["query_len"][:k]).sum())
        return cells, time.perf_counter() - t0

    k0 = min(n, max(cores * 32, 512))
    c0, t0 = run(k0)                                   # calibration (also warms the allocator and the thread pool)
    k = int(min(n, max(k0, k0 * seconds / max(t0, 1e-3))))
    cells, t = run(k)
    if gpu_scores is not None and not (scores[:k] == gpu_scores[:k]).all():
        raise SystemExit("bench: GPU tile scores differ from the CPU checker on the baseline sample")
    return {"value": cells / t / 1e9, "unit": UNIT, "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": f"{k} tiles of the same batch ({cells / 1e9:.2f} G cells) in {t:.1f} s, "
                      f"{'reference AlignWithBT (align.cpp:60-233) via oracle/_ref' if use_ref else 'oracle/gact_oracle.c: the same recurrence without the per-tile 16.8 MB vector<vector<int>> of align.cpp:85'}, "
                      f"OpenMP over tiles, {cores} threads", "seconds": t, "tiles": k}


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mb = make_batch(min(args.tiles, 1 << 16), 42)
    per = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    runs = [cpu_arm(args, mb, per) for _ in range(args.warmup + args.steps)]
    timed = runs[args.warmup:]
    cells = sum(r["value"] * r["seconds"] for r in timed)
    secs = sum(r["seconds"] for r in timed)
    v = cells / secs
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config(args), "sample_tiles_per_step": timed[-1]["tiles"],
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": timed[-1]["cores"], "kind": timed[-1]["kind"],
                             "sample": timed[-1]["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


OVERLAP_RE = None


def _run_darwin(exe, wd, ref, reads, threads, env=None, timeout=900):
    """Run a darwin binary in `wd`; returns (stdout, {query name -> [lines]} from darwin.*.out, wall seconds)."""
    import re
    global OVERLAP_RE
    if OVERLAP_RE is None:
        OVERLAP_RE = re.compile(r"query_id: (\S+),")
    for fn in os.listdir(wd):
        if fn.startswith("darwin.") and fn.endswith(".out"):
            os.remove(os.path.join(wd, fn))
    t0 = time.perf_counter()
    r = subprocess.run([exe, ref, reads, str(threads)], cwd=wd, capture_output=True, text=True,
                       env=dict(os.environ, **(env or {})), timeout=timeout)
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("%s exited %d: %s" % (os.path.basename(exe), r.returncode, r.stderr[-300:]))
    lines = []
    for fn in sorted(os.listdir(wd)):
        if fn.startswith("darwin.") and fn.endswith(".out"):
            lines += open(os.path.join(wd, fn)).read().splitlines()
    return r.stdout, lines, wall


def reads_leg(n_gpus, cpu_arm_reads=200):
    """Application-level leg of the metric (BASELINE.json: reads/s at 1/2/4/8 GPUs): the drop-in `darwin` binary
    (C++ host + engine) on the FULL config 3 / 4 workload -- 100 Mbp reference (20 x 5 Mbp, seed 3), 50 MB of PacBio-like
    ~10 kb reads (seed 4), params.cfg defaults, reads sharded over n_gpus GPUs (strong scaling) -- plus, for n_gpus > 1,
    a weak-scaled run (n_gpus x 50 MB of reads).  The reference CPU build (oracle/_ref/darwin_ref, all host cores) runs
    on the first `cpu_arm_reads` reads in the same run; its sorted|uniq output must equal ours on those reads."""
    import re
    import shutil
    import tempfile
    import synth
    exe = os.path.join(ROOT, "darwin-gpu_b200", "darwin")
    if not os.path.exists(exe):
        return {"unavailable": "darwin-gpu_b200/darwin not built"}
    cores = usable_cores()
    wd = tempfile.mkdtemp(prefix="bench_reads_")
    try:
        rng = np.random.default_rng(3)
        genome = [synth.random_genome(5_000_000, rng) for _ in range(20)]
        synth.write_fasta(os.path.join(wd, "ref.fasta"), [f"chr{i}" for i in range(20)], genome)
        names, reads = synth.sample_reads(genome, 50_000_000, np.random.default_rng(4), mean=10000, sd=3000, lo=1000, hi=30000)
        synth.write_fasta(os.path.join(wd, "reads.fasta"), names, reads)
        open(os.path.join(wd, "params.cfg"), "w").write(open(os.path.join(ROOT, "darwin-gpu_b200", "params.cfg")).read())
        n_sub = min(cpu_arm_reads, len(reads))
        synth.write_fasta(os.path.join(wd, "reads_sub.fasta"), names[:n_sub], reads[:n_sub])
        sub_names = set(names[:n_sub])

        def ours(reads_file, n_reads, label):
            out, lines, wall = _run_darwin(exe, wd, "ref.fasta", reads_file, cores, env={"DARWIN_GPUS": str(n_gpus)})
            summ = json.loads(re.search(r"DARWIN_B200_SUMMARY (\{.*\})", out).group(1))
            align_s = max(summ["align_phase_ms"], 1e-3) / 1e3
            sub = sorted(set(ln for ln in lines if OVERLAP_RE.search(ln) and OVERLAP_RE.search(ln).group(1) in sub_names))
            rec = {"workload": label, "gpus": summ["gpus"], "reads": n_reads, "reads_per_s": n_reads / align_s,
                   "align_phase_ms": summ["align_phase_ms"], "worker_setup_ms": summ["worker_setup_ms"],
                   "reads_per_s_incl_worker_setup": n_reads / (align_s + summ["worker_setup_ms"] / 1e3),
                   "wall_s": wall, "wall_s_in_process": summ["wall_s"], "gpu_init_ms": summ["gpu_init_ms"],
                   "teardown_ms": summ["teardown_ms"], "candidates": summ["candidates"], "tiles": summ["tiles"],
                   "cells": summ["cells"], "gcups_align_phase": summ["cells"] / align_s / 1e9,
                   "gact_kernel_ms": summ["gact_kernel_ms"], "chain_batches": summ["chain_batches"],
                   "overlap_lines": len(lines), "unique_overlap_lines": len(set(lines)), "host_threads": cores}
            return rec, sub

        strong, sub_strong = ours("reads.fasta", len(reads),
                                  "config 3/4 at full size: 50 MB PacBio-like ~10 kb reads (15 % error) vs 100 Mbp reference, "
                                  "params.cfg defaults; D-SOFT, GACT extension and seed table on the GPU(s)")
        strong["read_bases"] = int(sum(len(x) for x in reads))
        res = {"bracket": "align_phase_ms is the reference's 'seed table querying + aligning' bracket (darwin.cpp:615-639) "
                          "except that worker threads are started, bound to their device and their darwin.<tid>.out created "
                          "BEFORE it opens (the reference does that inside, darwin.cpp:174-175); worker_setup_ms is that "
                          "part, reads_per_s_incl_worker_setup the like-for-like figure; wall_s is the whole process",
               "strong": strong}
        # ---- reference CPU build on a subset, same run, parity on that subset ----
        ref_exe = os.path.join(ROOT, "oracle", "_ref", "darwin_ref")
        if os.path.exists(ref_exe):
            out, lines, wall = _run_darwin(ref_exe, wd, "ref.fasta", "reads_sub.fasta", cores, timeout=1200)
            m = re.search(r"Time elapsed \(seed table querying \+ aligning\): (\d+) msec", out)
            ref_align_s = max(int(m.group(1)), 1) / 1e3
            ref_lines = sorted(set(lines))
            res["reference_cpu"] = {"kind": "reference CPU build (darwin.cpp + gact.cpp + align.cpp compiled in place), its own "
                                            "'seed table querying + aligning' bracket", "reads": n_sub, "threads": cores,
                                    "align_phase_ms": ref_align_s * 1e3, "reads_per_s": n_sub / ref_align_s, "wall_s": wall,
                                    "unique_overlap_lines": len(ref_lines)}
            res["sorted_uniq_identical_on_subset"] = bool(ref_lines == sub_strong)
            res["reads_per_s_vs_reference_cpu"] = strong["reads_per_s"] / res["reference_cpu"]["reads_per_s"]
            if ref_lines != sub_strong:
                res["parity_error"] = "sorted|uniq output differs from the reference CPU build on the %d-read subset" % n_sub
        else:
            ref_lines = None
            res["reference_cpu"] = {"unavailable": "oracle/_ref/darwin_ref not built"}
        # ---- weak scaling: n_gpus x 50 MB ----
        if n_gpus > 1:
            # every GPU gets the same work as the single GPU of the strong-scaling run: the 50 MB read set replicated
            # n_gpus times (names prefixed), so per-GPU work is exactly fixed as N grows
            all_names, all_reads = list(names), list(reads)
            for r in range(1, n_gpus):
                all_names += [f"W{r}x{nm}" for nm in names]
                all_reads += reads
            # replicas are concatenated: the contiguous shards of ceil(N / G) reads are then one replica each
            synth.write_fasta(os.path.join(wd, "reads_weak.fasta"), all_names, all_reads)
            weak, sub_weak = ours("reads_weak.fasta", len(all_reads), f"weak scaling: the 50 MB read set replicated {n_gpus} x (one replica "
                                                                     "per GPU) vs the same 100 Mbp reference")
            res["weak"] = weak
            if ref_lines is not None:
                res["weak_sorted_uniq_identical_on_subset"] = bool(ref_lines == sub_weak)
                if ref_lines != sub_weak:
                    res["parity_error"] = "weak-scaled run: output differs from the reference CPU build on the subset"
        return res
    finally:
        shutil.rmtree(wd, ignore_errors=True)


PARAMS_TMPL = None


def params_cfg(tile=None, overlap=None):
    """params.cfg defaults of the package, optionally with another tile_size / tile_overlap."""
    import re
    txt = open(os.path.join(ROOT, "darwin-gpu_b200", "params.cfg")).read()
    if tile is not None:
        txt = re.sub(r"(?m)^tile_size\s*=.*$", f"tile_size = {tile}", txt)
        txt = re.sub(r"(?m)^tile_overlap\s*=.*$", f"tile_overlap = {overlap}", txt)
    return txt


def config_at_size(which, n_gpus, cpu_reads):
    """BASELINE.json configs 1 and 5 at their named sizes: one JSON record per run, the reference CPU build on
    a `cpu_reads`-read subset with its sorted|uniq output compared with ours, line count + md5 of our full output."""
    import hashlib
    import re
    import shutil
    import tempfile
    import synth
    exe = os.path.join(ROOT, "darwin-gpu_b200", "darwin")
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "darwin_ref")
    cores = usable_cores()
    wd = tempfile.mkdtemp(prefix=f"bench_config{which}_")
    out = {"config": which, "gpus": n_gpus, "host_threads": cores, "runs": []}
    try:
        if which == 1:
            # README:25 stand-in: 4 641 652 bp genome (seed 1), PBSIM-CLR-like reads to 10x, self-aligned, params.cfg defaults
            genome = [synth.random_genome(4_641_652, np.random.default_rng(1))]
            names, reads = synth.sample_reads(genome, 46_416_520, np.random.default_rng(2), mean=3000, sd=2300, lo=100, hi=25000)
            synth.write_fasta(os.path.join(wd, "reads.fasta"), names, reads)
            jobs = [("self-alignment", "reads.fasta", "reads.fasta", None, None)]
            out["workload"] = (f"config 1: {len(reads)} PBSIM-CLR-like reads ({sum(len(r) for r in reads)} bases = 10x of a 4 641 652 bp "
                               "genome, 15 % error), reads.fasta against itself, params.cfg defaults")
        else:
            rng = np.random.default_rng(3)
            genome = [synth.random_genome(5_000_000, rng) for _ in range(20)]
            synth.write_fasta(os.path.join(wd, "ref.fasta"), [f"chr{i}" for i in range(20)], genome)
            names, reads = synth.sample_reads(genome, 51_000_000, np.random.default_rng(5), mean=30000, sd=3000, lo=20000, hi=40000)
            synth.write_fasta(os.path.join(wd, "reads.fasta"), names, reads)
            jobs = [(f"tile_size {t} / overlap {o}", "ref.fasta", "reads.fasta", t, o) for t, o in ((256, 96), (512, 192), (1024, 384))]
            out["workload"] = (f"config 5: {len(reads)} ultra-long reads (~30 kb, {sum(len(r) for r in reads)} bases, 15 % error) vs the "
                               "100 Mbp config-3 reference, tile_size 256/512/1024 with tile_overlap = 0.375 tile_size")
        n_sub = min(cpu_reads, len(reads))
        synth.write_fasta(os.path.join(wd, "reads_sub.fasta"), names[:n_sub], reads[:n_sub])
        sub_names = set(names[:n_sub])
        for label, ref_file, reads_file, tile, overlap in jobs:
            open(os.path.join(wd, "params.cfg"), "w").write(params_cfg(tile, overlap))
            stdout, lines, wall = _run_darwin(exe, wd, ref_file, reads_file, cores, env={"DARWIN_GPUS": str(n_gpus)})
            summ = json.loads(re.search(r"DARWIN_B200_SUMMARY (\{.*\})", stdout).group(1))
            uniq = sorted(set(lines))
            align_s = max(summ["align_phase_ms"], 1e-3) / 1e3
            rec = {"run": label, "reads": len(reads), "reads_per_s": len(reads) / align_s, "align_phase_ms": summ["align_phase_ms"],
                   "wall_s": wall, "tiles": summ["tiles"], "cells": summ["cells"], "gcups_align_phase": summ["cells"] / align_s / 1e9,
                   "candidates": summ["candidates"], "unique_overlap_lines": len(uniq),
                   "md5_sorted_uniq": hashlib.md5("\n".join(uniq).encode()).hexdigest()}
            if os.path.exists(ref_exe):
                # same inputs for both builds: the subset as reads against the run's reference file (for config 1 the full
                # read set, i.e. a reads-vs-reference run of the subset -- the self-alignment of the whole set would take
                # the CPU build hours)
                so, sl, sw = _run_darwin(exe, wd, ref_file, "reads_sub.fasta", cores, env={"DARWIN_GPUS": str(n_gpus)})
                ours_sub = sorted(set(sl))
                ro, rl, rw = _run_darwin(ref_exe, wd, ref_file, "reads_sub.fasta", cores, timeout=3000)
                m = re.search(r"Time elapsed \(seed table querying \+ aligning\): (\d+) msec", ro)
                ref_align_s = max(int(m.group(1)), 1) / 1e3
                ssum = json.loads(re.search(r"DARWIN_B200_SUMMARY (\{.*\})", so).group(1))
                rec["subset"] = {"reads": n_sub, "reference_cpu_align_phase_ms": ref_align_s * 1e3,
                                 "reference_cpu_reads_per_s": n_sub / ref_align_s, "reference_cpu_wall_s": rw,
                                 "ours_align_phase_ms": ssum["align_phase_ms"], "unique_overlap_lines": len(ours_sub),
                                 "sorted_uniq_identical": bool(sorted(set(rl)) == ours_sub)}
                rec["reads_per_s_vs_reference_cpu"] = rec["reads_per_s"] / rec["subset"]["reference_cpu_reads_per_s"]
            out["runs"].append(rec)
        return out
    finally:
        shutil.rmtree(wd, ignore_errors=True)


def gpu_baseline_leg(n_tiles=1 << 16):
    """The reference's own GPU kernel on the same box (tools/ref_gpu_bench.py in a subprocess)."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_gpu_bench.py"), str(n_tiles), "42"],
                           capture_output=True, text=True, timeout=600)
        for ln in reversed(r.stdout.splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": "reference GPU run printed no result (rc %d): %s" % (r.returncode, (r.stderr or r.stdout)[-200:])}
    except Exception as ex:
        return {"unavailable": repr(ex)[:200]}


def shard_range(n_items, world, rank):
    """Contiguous shard of `n_items` for `rank` -- the reference's rule ceil(N / threads) per worker
    (darwin.cpp:620-623), also used by host/darwin_main.cpp for its per-GPU read ranges."""
    per = -(-n_items // max(world, 1))
    lo = min(n_items, per * rank)
    return lo, min(n_items, lo + per)


def reduce_over_ranks(dist, maxima, sums, device):
    """MAX-reduce the per-rank timings and SUM-reduce the per-rank work; identity for one rank.
    Works with NCCL (device tensors) and gloo (CPU tensors, used by the CPU tests)."""
    import torch
    t = torch.tensor(list(maxima), dtype=torch.float64, device=device)
    c = torch.tensor(list(sums), dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return t.tolist(), c.tolist()


# ---------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return reference_main(args)
    if args.config:
        print(json.dumps(config_at_size(args.config, args.gpus, args.cpu_reads)))
        return

    import torch
    import pygact as G

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device (there is no CPU fallback for the GACT path)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    mb = make_batch(args.tiles, 42 + rank)
    n = args.tiles
    cells = int((mb["ref_len"].astype(np.int64) * mb["query_len"]).sum())
    stream = torch.cuda.Stream(device=local)
    eng = G.GactEngine(*SCORES, tile_size=TILE, tile_overlap=OVERLAP, device=local, max_tiles=n,
                       stream=stream.cuda_stream)
    if args.kernel:
        eng.set_kernel(args.kernel)
    eng.upload(G.SET_REF, [mb["ref"].tobytes()])
    eng.upload(G.SET_READS, [mb["query"].tobytes()])
    d = G.make_descs(n)
    for k in ("ref_off", "query_off", "ref_len", "query_len", "reverse", "first"):
        d[k] = mb[k]
    d["ref_set"], d["query_set"] = G.SET_REF, G.SET_READS

    # ---- leg 1: device-resident (value) --------------------------------------------------
    eng.stage(d)
    sampler = ClockSampler(local)
    sampler.start()                    # nvidia-smi comes up during the warm-up; only samples after begin() are used
    for _ in range(args.warmup):
        eng.run_staged()
    eng.sync()
    barrier()
    sampler.begin()
    launches0 = eng.stats()["kernel_launches"]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for s in range(args.steps):
            eng.run_staged()
            ev[s + 1].record(stream)
    barrier()
    launches = eng.stats()["kernel_launches"] - launches0
    dev_ms = ev[0].elapsed_time(ev[-1])
    per_step_ms = [ev[s].elapsed_time(ev[s + 1]) for s in range(args.steps)]
    res_dev, st_dev = eng.fetch_staged()
    path_info = eng.tile_path_info()

    # ---- leg 2: end to end through the C ABI with host buffers (e2e) ----------------------
    chunk = min(args.chunk, n)
    # pipeline ramp: a quarter- and a half-size batch at both ends, so the first kernel starts after a short upload
    # and only a short download + copy-out is left after the last one
    sizes = [chunk // 4, chunk // 2] if n >= 4 * chunk and chunk >= 4096 else []
    mid = n - 2 * sum(sizes)
    sizes = sizes + [chunk] * (mid // chunk) + ([mid % chunk] if mid % chunk else []) + sizes[::-1]
    bounds, lo = [], 0
    for sz in sizes:
        bounds.append((lo, lo + sz))
        lo += sz
    assert lo == n

    e2e_res = np.zeros(n, dtype=G.TILE_RESULT_DTYPE)          # caller-owned host result buffers
    e2e_st = np.zeros((n, eng.pitch), dtype=np.uint32)

    def e2e_step():
        pend = []
        for lo, hi in bounds:
            eng.submit(d[lo:hi])
            pend.append((lo, hi))
            if len(pend) == G.MAX_INFLIGHT:
                a, b = pend.pop(0)
                eng.wait(e2e_res[a:b], e2e_st[a:b])
        while pend:
            a, b = pend.pop(0)
            eng.wait(e2e_res[a:b], e2e_st[a:b])
        return e2e_res

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res_e2e = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()            # sampled across both timed regions (device-resident and e2e legs)
    assert (res_e2e == res_dev).all(), "e2e and device-resident legs disagree"
    pitch = eng.pitch
    h2d = n * 36 + int(mb["first"].sum()) * 4
    d2h = n * (24 + pitch * 4)

    # ---- reductions over ranks ---------------------------------------------------------------
    (dev_ms_max, e2e_ms_max), (total_cells,) = reduce_over_ranks(dist, [dev_ms, e2e_s * 1e3], [float(cells)],
                                                                 f"cuda:{local}")

    if rank == 0:
        gcups = total_cells * args.steps / (dev_ms_max * 1e-3) / 1e9
        e2e_gcups = total_cells * args.steps / (e2e_ms_max * 1e-3) / 1e9
        # roofline: integer/DPX issue rate, measured live on this GPU
        peak_alu = G.int_peak(2, local)          # VIADDMNMX.S32: one ALU-pipe lane-op per thread-instruction
        peak_mix = G.int_peak(7, local)          # 1:1 ALU:FMA-pipe mix
        kernel_ms = float(np.mean(per_step_ms))
        gcups_rank0 = cells / (kernel_ms * 1e-3) / 1e9
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_bytes = (n * 32 + n * (24 + 4 * pitch) + cells * 2 * 0.25 / TILE)   # descs + results + 2-bit bases
        packed = eng.get_kernel() == 2
        floor_ops = ALU_OPS_PER_CELL_FLOOR if packed else 2 * ALU_OPS_PER_CELL_FLOOR
        achieved = gcups_rank0 * floor_ops
        # executed-instruction figures of the step's kernels from this round's ncu captures (tools/ncu_summary.py -> profiles/);
        # nothing from a profiler run is used as a bench value, the counters only turn the measured step time into pipe utilisation
        issue, traffic = None, None
        sm_clock_hz = (clocks.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) * 1e6
        if packed and TILE == 320:
            caps = {}
            for name, path in NCU_SUMMARIES:
                try:
                    caps[name] = json.load(open(path))
                except Exception:
                    pass
            n_it = path_info["inter_task"]
            cells_it = n_it * TILE * TILE
            # the engine's routing rule (check_descs): non-first, not full, query window of at most 160 columns -> narrow mapping
            rl64, ql64 = mb["ref_len"].astype(np.int64), mb["query_len"].astype(np.int64)
            is_narrow = (mb["first"] == 0) & (ql64 <= 160) & ~((rl64 == TILE) & (ql64 == TILE))
            cells_narrow = int((rl64 * ql64)[is_narrow].sum()) if os.environ.get("GACT_NARROW", "1") != "0" and "wavefront_narrow" in caps else 0
            share = {"inter_task": cells_it, "wavefront": cells - cells_it - cells_narrow, "wavefront_narrow": cells_narrow}
            if all(k in caps and caps[k].get("cells") for k, v in share.items() if v > 0):
                smsp = 4 * 148
                inst = sum(caps[k]["smsp__inst_executed.sum"] / caps[k]["cells"] * v for k, v in share.items() if v > 0)
                alu = sum(caps[k].get("alu_pipe_warp_inst", 0) / caps[k]["cells"] * v for k, v in share.items() if v > 0)
                fma = sum(caps[k].get("fma_pipe_warp_inst", 0) / caps[k]["cells"] * v for k, v in share.items() if v > 0)
                traffic = sum((caps[k]["dram__bytes_read.sum"] + caps[k]["dram__bytes_write.sum"]) / caps[k]["cells"] * v
                              for k, v in share.items() if v > 0)
                issue = {"kernels": {k: {"source": os.path.relpath(dict(NCU_SUMMARIES)[k], ROOT), "capture": caps[k].get("capture"),
                                         "cells_in_this_step": v, "warp_inst_per_cell": caps[k]["smsp__inst_executed.sum"] / caps[k]["cells"],
                                         "alu_pipe_busy_ncu_pct": caps[k].get("sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"),
                                         "issue_active_ncu_pct": caps[k].get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                         "dram_bytes_per_cell": (caps[k]["dram__bytes_read.sum"] + caps[k]["dram__bytes_write.sum"]) / caps[k]["cells"]}
                                     for k, v in share.items() if v > 0},
                         "warp_inst_per_step": inst, "alu_pipe_warp_inst_per_step": alu, "fma_pipe_warp_inst_per_step": fma,
                         # one warp instruction per cycle per SM sub-partition is the issue peak; the ALU pipe takes one every two cycles
                         "issue_slot_frac": inst / (kernel_ms * 1e-3 * smsp * sm_clock_hz),
                         "alu_pipe_frac": alu / (kernel_ms * 1e-3 * smsp * sm_clock_hz * 0.5),
                         "sm_clock_mhz_used": sm_clock_hz / 1e6,
                         "note": "the inter-task kernel overlaps the wavefront kernels on two streams; fractions are over the whole step"}
        roof = {"bound": "int_issue",
                "kernel": ("tile step = gact_tile_it_kernel (%d of %d tiles: full, non-first) overlapped with gact_first_s16h / "
                           "gact_tile_s16h_kernel<10,16,true> / <5,16,true> (the rest; the narrow mapping for query windows of at "
                           "most 160 columns), packed s16x2 DPX" % (path_info["inter_task"], n)) if packed
                else "gact_tile_i32 kernel",
                "achieved": achieved, "peak": peak_alu,
                "unit": "G ALU-pipe lane-ops/s: achieved = cells/s x %.1f (algorithmic floor of the %s recurrence, 5 ALU-pipe "
                        "instructions per cell pair: PRMT score select, 3 VIADDMNMX, 1 VIMNMX3); peak = measured ALU-pipe issue rate "
                        "(gact_int_peak, VIADDMNMX, this run)" % (floor_ops, "packed s16x2" if packed else "int32"),
                "frac": achieved / peak_alu,
                "traffic": traffic,
                "traffic_note": "DRAM bytes per step (dram__bytes_read.sum + dram__bytes_write.sum of the step's kernels from this round's "
                                "ncu --set full captures, scaled by the cells each kernel got): the inter-task kernel's strip edges "
                                "(16 B per row, lane and strip) and the wavefront kernel's direction windows; algorithmic traffic is "
                                "0.3 KB per tile",
                "alu_ops_per_cell_floor": floor_ops, "lane_width": "s16x2" if packed else "s32",
                "gcups_at_floor_roofline": peak_alu / floor_ops,
                "executed": issue,
                "survey_model": {"ops_per_cell": OPS_PER_CELL_SURVEY, "gcups_roofline_int32_alu": peak_alu / OPS_PER_CELL_SURVEY,
                                 "frac_of_int32_roofline": gcups_rank0 * OPS_PER_CELL_SURVEY / peak_alu,
                                 "note": "SURVEY 8d's 17 scalar int32 instructions per cell (round 1's denominator): the packed tagged-max "
                                         "kernel needs far fewer, so this fraction exceeds 1 and is kept only for continuity"},
                "peak_alu_fma_mix": peak_mix, "peak_source": "gact_int_peak (own microbenchmark, measured in this run)",
                "kernel_ms": kernel_ms,
                "hbm": {"achieved": hbm_bytes / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": hbm_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
        line = {"metric": METRIC, "value": gcups, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "s16x2" if eng.get_kernel() == 2 else "int32",
                "data": "synthetic", "config": config(args), "kernel_variant": eng.get_kernel(),
                "gcups_per_gpu": gcups / world,
                "tiles_per_s": n * world * args.steps / (dev_ms_max * 1e-3),
                "e2e": {"value": e2e_gcups, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms_max / args.steps, "api": "gact_engine_submit/wait, %d batches of up to %d tiles (quarter/half-size batches at both ends), "
                               "%d in flight" % (len(bounds), chunk, G.MAX_INFLIGHT)},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
                "tile_routing": {"inter_task_kernel_tiles": path_info["inter_task"], "handed_back_to_wavefront_kernel": path_info["handed_back"],
                                 "wavefront_kernel_tiles": n - path_info["inter_task"]}}
        if world == 1 and not args.no_cpu_baseline:
            small = {k: (v[:1 << 16] if k not in ("ref", "query") else v) for k, v in mb.items()}
            keep = ("value", "unit", "cores", "kind", "sample")
            line["cpu_baseline"] = {k: v for k, v in cpu_arm(args, small, args.cpu_seconds, res_dev["score"]).items() if k in keep}
            # the same recurrence without the reference's per-tile 16.8 MB allocation (align.cpp:85): the fair CPU datapoint
            line["cpu_baseline_port"] = {k: v for k, v in cpu_arm(args, small, max(4.0, args.cpu_seconds / 2), res_dev["score"],
                                                                  force_port=True).items() if k in keep}
    else:
        line = None
    eng.close()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.cuda.synchronize()
        line["gpu_baseline"] = gpu_baseline_leg()
        gb = line["gpu_baseline"]
        if gb.get("value"):
            line["speedup_vs_reference_gpu_kernel"] = {"device_resident": line["value"] / gb["value"],
                                                       "end_to_end": line["e2e"]["value"] / gb["e2e_value"]}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()       # the application leg below runs without NCCL: the other ranks are gone by then
    if rank == 0:
        if not args.no_reads_leg:
            try:
                line["reads"] = reads_leg(world)
            except Exception as ex:          # the application leg must never cost the kernel numbers
                line["reads"] = {"unavailable": repr(ex)[:200]}
        print(json.dumps(line))


if __name__ == "__main__":
    main()
