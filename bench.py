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
  roofline / cpu_baseline / clocks: see DESIGN.md.

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
NCU_DRAM_BYTES_PER_TILE = (705.809152e6 + 2.405678e9) / 131072     # see roofline.traffic_note
OPS_PER_CELL = 17.0          # scalar int32 instructions per DP cell
TILE, OVERLAP = 320, 120
SCORES = (1, -1, -1, -1)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tiles", type=int, default=1 << 20, help="tiles per GPU per step")
    ap.add_argument("--chunk", type=int, default=1 << 16, help="tiles per submit() in the e2e leg")
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 int32, 2 s16x2")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reads-leg", action="store_true", help="skip the application-level reads/s leg")
    return ap.parse_args()


def config(args, extra=None):
    c = {"workload": "config2: tile microbatch, 1Mi tile pairs tile_size=320 overlap=120, 15% PacBio-like error, "
                     "82% full/18% edge, 5.5% first, seed 42 (+rank)",
         "tiles_per_gpu_per_step": args.tiles, "tile_size": TILE, "tile_overlap": OVERLAP,
         "scores": list(SCORES), "parallelism": f"reads/tiles sharded over {args.gpus} GPU(s), no collective",
         "l2_policy": "no explicit flush: at the default size one step touches more than the 126 MB L2 (descriptors + "
                      "tile order 36 MiB, packed bases of 1 Mi tile windows, results + states 128 MiB, direction-window "
                      "scratch 115 MB), and the kernel is ALU-bound at 0.003 algorithmic bytes per cell"}
    if extra:
        c.update(extra)
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
        self.lines = []

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
            self.lines.append(ln.strip())

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
        for ln in self.lines:
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


def cpu_arm(args, mb, seconds, gpu_scores=None):
    """Reference CPU AlignWithBT (oracle/_ref) or the oracle port, all host threads, bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    cores = usable_cores()
    n = len(mb["ref_off"])
    od = np.zeros(n, dtype=O.TILE_DESC_DTYPE)
    for k in ("ref_off", "query_off", "ref_len", "query_len", "reverse", "first"):
        od[k] = mb[k]
    use_ref = O.ref_available()

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

    k0 = min(n, max(cores * 4, 64))
    c0, t0 = run(k0)                                   # calibration (also warms the allocator)
    k = int(min(n, max(k0, k0 * seconds / max(t0, 1e-3))))
    cells, t = run(k)
    if gpu_scores is not None and not (scores[:k] == gpu_scores[:k]).all():
        raise SystemExit("bench: GPU tile scores differ from the CPU checker on the baseline sample")
    return {"value": cells / t / 1e9, "unit": UNIT, "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": f"{k} tiles of the same batch ({cells / 1e9:.2f} G cells) in {t:.1f} s, "
                      f"{'reference AlignWithBT (align.cpp:60-233) via oracle/_ref' if use_ref else 'oracle C port'}, "
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
            "config": config(args, {"sample_tiles_per_step": timed[-1]["tiles"]}),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": timed[-1]["cores"], "kind": timed[-1]["kind"],
                             "sample": timed[-1]["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def reads_leg(n_gpus):
    """Application-level leg of the metric: reads/s of the drop-in `darwin` binary (C++ host + engine) in the
    reference's own timed bracket ("seed table querying + aligning", darwin.cpp:615-639) on a bounded config-3-shaped
    workload: 25 MB of PacBio-like ~10 kb reads against a 20 Mbp reference, reads sharded over n_gpus GPUs."""
    import re
    import tempfile
    import synth
    exe = os.path.join(ROOT, "darwin-gpu_b200", "darwin")
    if not os.path.exists(exe):
        return {"unavailable": "darwin-gpu_b200/darwin not built"}
    wd = tempfile.mkdtemp(prefix="bench_reads_")
    rng = np.random.default_rng(3)
    genome = [synth.random_genome(1000000, rng) for _ in range(20)]
    synth.write_fasta(os.path.join(wd, "ref.fasta"), [f"chr{i}" for i in range(20)], genome)
    names, reads = synth.sample_reads(genome, 25_000_000, np.random.default_rng(4), mean=10000, sd=3000, lo=1000, hi=30000)
    synth.write_fasta(os.path.join(wd, "reads.fasta"), names, reads)
    open(os.path.join(wd, "params.cfg"), "w").write(open(os.path.join(ROOT, "darwin-gpu_b200", "params.cfg")).read())
    r = subprocess.run([exe, "ref.fasta", "reads.fasta", str(usable_cores())], cwd=wd, capture_output=True, text=True,
                       env=dict(os.environ, DARWIN_GPUS=str(n_gpus)), timeout=600)
    if r.returncode != 0:
        return {"unavailable": "darwin exited %d: %s" % (r.returncode, r.stderr[-200:])}
    summ = json.loads(re.search(r"DARWIN_B200_SUMMARY (\{.*\})", r.stdout).group(1))
    lines = 0
    for fn in os.listdir(wd):
        if fn.startswith("darwin.") and fn.endswith(".out"):
            lines += sum(1 for _ in open(os.path.join(wd, fn)))
    align_s = max(summ["align_phase_ms"], 1) / 1e3
    return {"reads_per_s": len(reads) / align_s, "align_phase_ms": summ["align_phase_ms"], "reads": len(reads),
            "read_bases": int(sum(len(x) for x in reads)), "gpus": n_gpus, "tiles": summ["tiles"], "cells": summ["cells"],
            "gcups_align_phase": summ["cells"] / align_s / 1e9, "overlap_lines": lines,
            "workload": "config3-shaped, bounded: 25 MB PacBio-like ~10 kb reads vs 20 x 1 Mbp reference, params.cfg "
                        "defaults, D-SOFT + GACT extension on the GPU, timed bracket = the reference's "
                        "'seed table querying + aligning'"}


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
    for _ in range(args.warmup):
        eng.run_staged()
    eng.sync()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
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
        achieved = gcups_rank0 * OPS_PER_CELL
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_bytes = (n * 32 + n * (24 + 4 * pitch) + cells * 2 * 0.25 / TILE)   # descs + results + 2-bit bases
        packed = eng.get_kernel() == 2
        width = 2.0 if packed else 1.0         # cells per lane-op: the packed kernel computes two s16 cells per 32-bit lane
        roof = {"bound": "int_issue", "kernel": "gact_tile_s16h/s16 kernel (packed s16x2 DPX)" if packed else "gact_tile_i32 kernel",
                "achieved": achieved, "peak": peak_alu * width,
                "unit": "G algorithmic int ops/s (17 per DP cell, SURVEY 8d); peak = measured ALU-pipe lane-op rate x cells per lane-op",
                "frac": achieved / (peak_alu * width),
                "traffic": NCU_DRAM_BYTES_PER_TILE * n if (packed and TILE == 320) else None,
                "traffic_note": "DRAM bytes per launch = ncu --set full capture of this kernel (131072-tile launch: "
                                "dram__bytes_read.sum 706 MB + dram__bytes_write.sum 2406 MB, profiles/r1_s16h_tile_kernel_ncu.txt) "
                                "scaled by tile count; it is the per-warp direction window (22 KB/tile, written once, read by the "
                                "traceback) spilling from L2, about 0.7 TB/s = 10 % of HBM bandwidth, not the bound",
                "ops_per_cell": OPS_PER_CELL, "lane_width": "s16x2" if packed else "s32",
                "peak_alu_lane_ops": peak_alu, "frac_of_int32_roofline": achieved / peak_alu,
                "gcups_roofline_int32_alu": peak_alu / OPS_PER_CELL, "gcups_roofline_s16x2_alu": 2 * peak_alu / OPS_PER_CELL,
                "note": "frac can exceed 1: the tagged-max formulation executes ~6 ALU-pipe instructions per cell instead "
                        "of the 17 (8.5 packed) the roofline model assumes; ncu of the same kernel: ALU pipe 86 % busy, issue slots 67 % "
                        "(profiles/r1_s16h_tile_kernel_ncu.txt)",
                "alu_pipe_busy_ncu": 0.8607 if (packed and TILE == 320) else None,   # sm__inst_executed_pipe_alu, profiles/r1_s16h_tile_kernel_ncu.txt
                "peak_alu_fma_mix": peak_mix, "peak_source": "gact_int_peak (own microbenchmark, measured in this run)",
                "kernel_ms": kernel_ms,
                "hbm": {"achieved": hbm_bytes / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": hbm_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
        line = {"metric": METRIC, "value": gcups, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "s16x2" if eng.get_kernel() == 2 else "int32",
                "data": "synthetic", "config": config(args, {"kernel_variant": eng.get_kernel()}),
                "gcups_per_gpu": gcups / world,
                "tiles_per_s": n * world * args.steps / (dev_ms_max * 1e-3),
                "e2e": {"value": e2e_gcups, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms_max / args.steps, "api": "gact_engine_submit/wait, %d batches of up to %d tiles (quarter/half-size batches at both ends), "
                               "%d in flight" % (len(bounds), chunk, G.MAX_INFLIGHT)},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof}
        if world == 1 and not args.no_cpu_baseline:
            small = {k: (v[:1 << 15] if k not in ("ref", "query") else v) for k, v in mb.items()}
            line["cpu_baseline"] = {k: v for k, v in cpu_arm(args, small, args.cpu_seconds, res_dev["score"]).items()
                                    if k in ("value", "unit", "cores", "kind", "sample")}
    else:
        line = None
    eng.close()
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
