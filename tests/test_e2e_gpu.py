"""GPU end-to-end parity: the drop-in `darwin` binary (C++ host + CUDA engine) against the
sorted|uniq output of the reference CPU build (golden fixtures produced by tests/golden/make_golden.py
with the unmodified reference, README:32 recipe)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "e2e_small")
EXE = os.path.join(ROOT, "darwin-gpu_b200", "darwin")

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


def run_darwin(workdir, ref, reads, threads, cfg, env=None, extra=()):
    os.makedirs(workdir, exist_ok=True)
    with open(os.path.join(workdir, "params.cfg"), "w") as f:
        f.write(PARAMS.format(**cfg))
    for fn in os.listdir(workdir):
        if fn.startswith("darwin.") and fn.endswith(".out"):
            os.remove(os.path.join(workdir, fn))
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([EXE, ref, reads, str(threads), *extra], cwd=workdir, capture_output=True, text=True, env=e, timeout=90)
    assert r.returncode == 0, r.stderr + r.stdout
    lines = []
    for fn in sorted(os.listdir(workdir)):
        if fn.startswith("darwin.") and fn.endswith(".out"):
            lines += open(os.path.join(workdir, fn)).read().splitlines()
    return sorted(set(lines)), r.stdout


def expected(tag):
    return open(os.path.join(GOLD, f"expected_{tag}.txt")).read().splitlines()


CFGS = {"t320": dict(ma=1, mi=-1, go=-1, ge=-1, ts=320, to=120),
        "t256": dict(ma=1, mi=-1, go=-1, ge=-1, ts=256, to=96),
        "t512_s2": dict(ma=2, mi=-3, go=-5, ge=-2, ts=512, to=192)}


@pytest.mark.parametrize("tag", sorted(CFGS))
@pytest.mark.parametrize("mode", ["default", "int32", "host_sched", "host_dsoft", "host_table"])
def test_reads_vs_reference_matches_cpu_build(tmp_path, tag, mode):
    env = {"default": {}, "int32": {"DARWIN_KERNEL": "1"}, "host_sched": {"DARWIN_CHAINS": "0"},
           "host_dsoft": {"DARWIN_DSOFT": "host", "DARWIN_CHAINS": "0"}, "host_table": {"DARWIN_SEEDTABLE": "host"}}[mode]
    got, out = run_darwin(str(tmp_path), os.path.join(GOLD, "ref.fasta"), os.path.join(GOLD, "reads.fasta"), 4,
                          CFGS[tag], env=env)
    assert got == expected(tag)
    assert "num_candidates:" in out and "Time elapsed (seed table querying + aligning)" in out


def test_self_alignment_same_file_suppression(tmp_path):
    reads = os.path.join(GOLD, "reads.fasta")
    got, _ = run_darwin(str(tmp_path), reads, reads, 3, CFGS["t320"], extra=("32", "64"))
    assert got == expected("self_t320")


def test_output_independent_of_thread_count(tmp_path):
    a, _ = run_darwin(str(tmp_path / "a"), os.path.join(GOLD, "ref.fasta"), os.path.join(GOLD, "reads.fasta"), 1, CFGS["t320"])
    b, _ = run_darwin(str(tmp_path / "b"), os.path.join(GOLD, "ref.fasta"), os.path.join(GOLD, "reads.fasta"), 7, CFGS["t320"])
    assert a == b == expected("t320")


def test_reference_gpu_host_code_links_against_library(tmp_path):
    """INTEGRATION.md route B: the reference's own darwin.cpp + gact.cpp (-D GPU), built in place and
    linked to libgact_b200.so through host/legacy_adapter.cpp, reproduces the CPU build's output.
    One host thread: the reference's GPU-build base conversion races with more."""
    exe = os.path.join(ROOT, "oracle", "_ref", "darwin_gpu_adapter")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/darwin_gpu_adapter not built (needs /root/reference at build time)")
    wd = str(tmp_path)
    with open(os.path.join(wd, "params.cfg"), "w") as f:
        f.write(PARAMS.format(**CFGS["t320"]))
    r = subprocess.run([exe, os.path.join(GOLD, "ref.fasta"), os.path.join(GOLD, "reads.fasta"), "1", "8", "64"],
                       cwd=wd, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = []
    for fn in sorted(os.listdir(wd)):
        if fn.startswith("darwin.") and fn.endswith(".out"):
            lines += open(os.path.join(wd, fn)).read().splitlines()
    assert sorted(set(lines)) == expected("t320")


ACGT = os.path.join(ROOT, "tests", "golden", "e2e_acgt")
ACGT_CFGS = {"t320": dict(ma=1, mi=-1, go=-1, ge=-1, ts=320, to=120),
             "t512": dict(ma=1, mi=-1, go=-1, ge=-1, ts=512, to=192),
             "t1024": dict(ma=1, mi=-1, go=-1, ge=-1, ts=1024, to=384),
             "t200_s2": dict(ma=2, mi=-3, go=-5, ge=-2, ts=200, to=60)}


@pytest.mark.parametrize("tag", sorted(ACGT_CFGS))
@pytest.mark.parametrize("mode", ["chains", "host_sched"])
def test_acgt_reads_device_chains_match_cpu_build(tmp_path, tag, mode):
    """ACGT-only data: sets are 2-bit packed, so the default path is D-SOFT + whole candidate extensions on the
    GPU (gact_engine_extend).  Both that path and the tile-round-trip scheduler must reproduce the CPU build."""
    env = {} if mode == "chains" else {"DARWIN_CHAINS": "0"}
    got, out = run_darwin(str(tmp_path), os.path.join(ACGT, "ref.fasta"), os.path.join(ACGT, "reads.fasta"), 4,
                          ACGT_CFGS[tag], env=env)
    exp = open(os.path.join(ACGT, f"expected_{tag}.txt")).read().splitlines()
    assert got == exp
    import json
    import re
    summ = json.loads(re.search(r"DARWIN_B200_SUMMARY (\{.*\})", out).group(1))
    assert summ["tiles"] > 0 and summ["cells"] > 0


def _gpu_count():
    import pygact
    return pygact.device_count()


@pytest.mark.parametrize("multiproc", ["1", "0"])
@pytest.mark.parametrize("gpus", [2, 4, 8])
def test_multi_gpu_sharded_output_matches_cpu_build(tmp_path, gpus, multiproc):
    """Reads sharded over several GPUs (DARWIN_GPUS; one worker process per GPU by default, one host thread per GPU with
    DARWIN_MULTIPROC=0; one engine and one darwin.<tid>.out per GPU either way): the concatenated, sorted|uniq output is
    the reference CPU build's, whatever the shard count (config 4)."""
    if _gpu_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    for gold, cfgs, tag in ((ACGT, ACGT_CFGS, "t320"), (GOLD, CFGS, "t320")):
        got, out = run_darwin(str(tmp_path / f"{os.path.basename(gold)}_{gpus}"), os.path.join(gold, "ref.fasta"),
                              os.path.join(gold, "reads.fasta"), 4, cfgs[tag],
                              env={"DARWIN_GPUS": str(gpus), "DARWIN_MULTIPROC": multiproc, "DARWIN_SORTED_OUT": "out.darwin"})
        assert open(os.path.join(str(tmp_path / f"{os.path.basename(gold)}_{gpus}"), "out.darwin")).read().splitlines() == got
        assert got == open(os.path.join(gold, f"expected_{tag}.txt")).read().splitlines()
        assert f"Using GPU: {gpus} device(s)" in out
        files = [fn for fn in os.listdir(str(tmp_path / f"{os.path.basename(gold)}_{gpus}")) if fn.startswith("darwin.") and fn.endswith(".out")]
        assert len(files) == gpus


@pytest.mark.parametrize("batch_reads", ["0", "7", "64"])
def test_pipelined_batches_and_sorted_unique_writer(tmp_path, batch_reads):
    """The shard's chains in batches of consecutive reads (DARWIN_BATCH_READS) give the same lines in the same order as one
    batch, and DARWIN_SORTED_OUT writes the README:32 `sort | uniq` file directly."""
    wd = str(tmp_path)
    got, out = run_darwin(wd, os.path.join(ACGT, "ref.fasta"), os.path.join(ACGT, "reads.fasta"), 4, ACGT_CFGS["t320"],
                          env={"DARWIN_BATCH_READS": batch_reads, "DARWIN_SORTED_OUT": "out.darwin"})
    exp = open(os.path.join(ACGT, "expected_t320.txt")).read().splitlines()
    assert got == exp
    assert open(os.path.join(wd, "out.darwin")).read().splitlines() == exp
    raw = open(os.path.join(wd, "darwin.0.out")).read().splitlines()
    one, _ = run_darwin(str(tmp_path / "one"), os.path.join(ACGT, "ref.fasta"), os.path.join(ACGT, "reads.fasta"), 4,
                        ACGT_CFGS["t320"], env={"DARWIN_BATCH_READS": "0"})
    assert raw == open(os.path.join(str(tmp_path / "one"), "darwin.0.out")).read().splitlines()
    import json
    import re
    summ = json.loads(re.search(r"DARWIN_B200_SUMMARY (\{.*\})", out).group(1))
    assert summ["align_phase_ms"] > 0 and summ["wall_s"] > 0 and summ["sorted_unique_lines"] == len(exp)
    assert "DARWIN_B200_TIMELINE" in out
