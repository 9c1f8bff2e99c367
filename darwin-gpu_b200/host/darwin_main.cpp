// darwin_main.cpp -- drop-in `darwin` command line on top of the B200 GACT engine.
//
//   ./darwin <REFERENCE>.fasta <READS>.fasta CPU_THREADS [NUM_BLOCKS THREADS_PER_BLOCK]
//
// Same surface as the reference (darwin.cpp:451-646, README:17-22): params.cfg in the
// working directory, positional arguments, `darwin.<tid>.out` overlap files, the
// "num_candidates: F R" and phase-time lines on stdout.  NUM_BLOCKS / THREADS_PER_BLOCK
// are accepted and ignored (the engine sizes its own persistent grid).  Additional
// controls come from the environment so that the reference form keeps working:
//   DARWIN_GPUS=<n>      GPUs to use (default: all visible); reads are sharded contiguously
//                        by ceil(num_reads / n), one host scheduler thread + engine per GPU
//   DARWIN_KERNEL=<0|1|2> kernel variant (auto / int32 / s16x2)
//   DARWIN_CHAINS=<1|0>   1 (default): whole candidate extensions on the device (gact_engine_extend) when the
//                        engine supports it; 0: tile-by-tile host scheduler (GactScheduler)
//   DARWIN_DSOFT=<gpu|host> where the D-SOFT filter runs (default gpu: gact_dsoft_run on the shard's GPU;
//                        host: SeedTable::dsoft on CPU_THREADS / n host threads)
//   DARWIN_MULTIPROC=<1|0> with more than one GPU: 1 (default) one worker PROCESS per GPU -- CUDA driver start-up and context
//                        creation are serialised inside a process (about 0.4 s per GPU) but run in parallel across processes,
//                        so the whole-program wall time no longer grows with the GPU count; 0: one host thread per GPU in
//                        this process (the reference's model, darwin.cpp:619-629).  Same files, same bracket, same output.
//   DARWIN_BATCH_READS=<n> reads per pipeline batch of a shard (default: auto, about 1 200; 0 = one batch): D-SOFT of
//                        batch k+1 and the output of batch k-1 overlap the alignment chains of batch k
//   DARWIN_SORTED_OUT=<file> additionally write the sorted, duplicate-free union of all darwin.<tid>.out lines
//                        (what README:32 builds with `cat darwin.*.out | sort | uniq`)
//
// Flow per GPU shard: D-SOFT on the host for every read of the shard (CPU_THREADS / n
// threads), then all candidates of the shard go through GactScheduler, then the overlap
// lines are written in the reference CPU build's order (per read: forward, then reverse).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <condition_variable>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <poll.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include "../../include/gact_b200.h"
#include "darwin_config.h"
#include "fasta_io.h"
#include "gact_scheduler.h"
#include "seed_table.h"

using namespace darwin;
using Clock = std::chrono::steady_clock;

static long ms_since(Clock::time_point t0)
{
    return (long)(std::chrono::duration<double, std::milli>(Clock::now() - t0).count() + 0.5);
}
static double us_since(Clock::time_point t0)
{
    return std::chrono::duration<double, std::micro>(Clock::now() - t0).count();
}

static std::mutex io_lock;

struct Shard {
    int tid = 0, device = 0;
    size_t first_read = 0, last_read = 0;     // [first, last)
    int dsoft_threads = 1;
    SchedulerStats stats;
    long dsoft_ms = 0, gact_ms = 0, init_ms = 0;
    double setup_us = 0;                       // worker start: thread creation, device binding, output file creation
    double dsoft_wait_us = 0, extend_wait_us = 0, write_us = 0;   // pipelined path: host time blocked on the device / writing
    int batches = 0;
    std::vector<std::string> text;             // output of the shard, one string per batch (for DARWIN_SORTED_OUT)
    std::vector<uint8_t> read_exc;             // per read of the shard: holds a byte other than ACGT (tile-by-tile path)
    uint64_t cand_fwd = 0, cand_rev = 0;
    std::string error;
    gact_engine *eng = nullptr;               // created before the align phase (like GPU_init, darwin.cpp:611)
    gact_dsoft *dsoft = nullptr;              // device-side D-SOFT filter (optional)
    gact_seed_table *seed_table = nullptr;    // seed-position table built on this shard's GPU (optional)
    double table_kernel_ms = 0.0;
};

// Lines that belong to one shard ("num_candidates: F R", "Time finding seeds", ...).  A worker process keeps its general
// chatter to itself (stdout -> /dev/null) and writes only these, whole lines at a time, to the real stdout.
static int g_shard_fd = -1;
static void shard_print(const std::string &line)
{
    std::lock_guard<std::mutex> lk(io_lock);
    if (g_shard_fd >= 0) {
        const std::string l = line + "\n";
        ssize_t r = write(g_shard_fd, l.data(), l.size());
        (void)r;
    } else {
        std::cout << line << std::endl;
    }
}

// One worker process per GPU (DARWIN_MULTIPROC): the coordinator side.  Workers are this binary again with
// DARWIN_WORKER="<index> <count>"; fd 3 carries the "go" byte to the worker, fd 4 its READY / DONE lines back.
struct WorkerProc {
    pid_t pid = -1;
    int ctl = -1, rep = -1;          // coordinator's ends
    std::string buf, ready, done;    // JSON payloads of the READY / DONE lines
    bool eof = false;
};

static double json_num(const std::string &js, const std::string &key, double dflt = 0.0)
{
    const std::string k = "\"" + key + "\": ";
    const size_t p = js.find(k);
    if (p == std::string::npos) return dflt;
    return atof(js.c_str() + p + k.size());
}

static int coordinate(int argc, char **argv, int G, const Params &cfg, int num_threads, bool same_file)
{
    const auto t_prog = Clock::now();
    std::vector<std::pair<std::string, double>> timeline;
    auto mark = [&](const char *what) { timeline.emplace_back(what, us_since(t_prog) / 1e3); };
    printf("CPU threads: %d\n", num_threads);
    printf("Scores: match = %d, mismatch = %d, gap_open = %d, gap_extend = %d\n", cfg.match, cfg.mismatch, cfg.gap_open, cfg.gap_extend);
    printf("Minimizer window size: %d\n", (int)cfg.window_size);
    printf("Using GPU: %d device(s), one worker process per device\n", G);
    fflush(stdout);
    std::vector<WorkerProc> w((size_t)G);
    for (int g = 0; g < G; g++) {
        int c2w[2], w2c[2];
        if (pipe2(c2w, O_CLOEXEC) || pipe2(w2c, O_CLOEXEC)) { perror("pipe2"); return 3; }
        const pid_t pid = fork();
        if (pid < 0) { perror("fork"); return 3; }
        if (pid == 0) {
            // keep clear of the fds we are about to overwrite
            const int rd = fcntl(c2w[0], F_DUPFD, 10), wr = fcntl(w2c[1], F_DUPFD, 10);
            dup2(rd, 3);
            dup2(wr, 4);
            const std::string env = std::to_string(g) + " " + std::to_string(G);
            setenv("DARWIN_WORKER", env.c_str(), 1);
            execv("/proc/self/exe", argv);
            perror("execv");
            _exit(127);
        }
        close(c2w[0]); close(w2c[1]);
        w[(size_t)g].pid = pid; w[(size_t)g].ctl = c2w[1]; w[(size_t)g].rep = w2c[0];
    }
    mark("workers_spawned");
    // read lines from every worker until `want` of them have delivered the given kind of line
    auto pump = [&](bool want_done) -> bool {
        for (;;) {
            size_t have = 0;
            for (auto &x : w) have += want_done ? !x.done.empty() : !x.ready.empty();
            if (have == w.size()) return true;
            std::vector<pollfd> pf;
            for (auto &x : w) if (!x.eof) pf.push_back(pollfd{x.rep, POLLIN, 0});
            if (pf.empty()) return false;
            if (poll(pf.data(), pf.size(), -1) < 0) { if (errno == EINTR) continue; return false; }
            for (auto &x : w) {
                if (x.eof) continue;
                for (auto &q : pf) if (q.fd == x.rep && (q.revents & (POLLIN | POLLHUP))) {
                    char tmp[4096];
                    const ssize_t r = read(x.rep, tmp, sizeof(tmp));
                    if (r <= 0) { x.eof = true; break; }
                    x.buf.append(tmp, (size_t)r);
                    size_t nl;
                    while ((nl = x.buf.find('\n')) != std::string::npos) {
                        const std::string line = x.buf.substr(0, nl);
                        x.buf.erase(0, nl + 1);
                        if (line.compare(0, 6, "READY ") == 0) x.ready = line.substr(6);
                        else if (line.compare(0, 5, "DONE ") == 0) x.done = line.substr(5);
                    }
                }
                if (x.eof && (want_done ? x.done.empty() : x.ready.empty())) return false;     // died before reporting
            }
        }
    };
    auto fail_all = [&](const char *what) {
        fprintf(stderr, "darwin: %s\n", what);
        for (auto &x : w) if (x.pid > 0) kill(x.pid, SIGTERM);
        for (auto &x : w) if (x.pid > 0) waitpid(x.pid, nullptr, 0);
        return 3;
    };
    if (!pump(false)) return fail_all("a worker process ended before it was ready (see its messages above)");
    mark("workers_ready_bracket_opens");
    const std::string &r0 = w[0].ready;
    std::cout << "\nLoading reference genome ...\nReference length: " << (long long)json_num(r0, "reference_length") << ", "
              << (int)json_num(r0, "ref_pieces") << " pieces\nTime elapsed (loading reference genome): "
              << (long)json_num(r0, "load_ref_ms") << " msec\n\nLoading reads ...\nNumber of reads: " << (long long)json_num(r0, "reads")
              << "\nTime elapsed (loading reads): " << (long)json_num(r0, "load_reads_ms") << " msec\n";
    double init_ms = 0, table_ms = 0, setup_ms = 0;
    for (int g = 0; g < G; g++) {
        init_ms = std::max(init_ms, json_num(w[(size_t)g].ready, "gpu_init_ms"));
        table_ms = std::max(table_ms, json_num(w[(size_t)g].ready, "seed_table_phase_ms"));
        setup_ms = std::max(setup_ms, json_num(w[(size_t)g].ready, "worker_setup_ms"));
        std::cout << "GPU " << g << " worker: CUDA up after " << (long)json_num(w[(size_t)g].ready, "cuda_up_ms") << " msec, init (context, engine, "
                  << "sequence upload) " << (long)json_num(w[(size_t)g].ready, "gpu_init_ms") << " msec, seed table build (H2D + kernels) "
                  << json_num(w[(size_t)g].ready, "seed_table_ms") << " msec, ready after " << (long)json_num(w[(size_t)g].ready, "ready_ms") << " msec\n";
    }
    std::cout << "Time elapsed (GPU init): " << (long)init_ms << " msec\nTime elapsed (seed position table construction): " << (long)table_ms
              << " msec\n\nFinding candidate bin locations for each read: \n" << G << " threads created\nSynchronizing all threads...\n" << std::flush;
    // ---- the timed bracket: from "every worker is ready" to "every worker has written its output" ----
    const auto t0 = Clock::now();
    for (auto &x : w) { const char go = 'G'; ssize_t r = write(x.ctl, &go, 1); (void)r; }
    if (!pump(true)) return fail_all("a worker process ended before it finished its shard");
    const double align_us = us_since(t0);
    mark("bracket_closes");
    std::cout << "Time elapsed (seed table querying + aligning): " << (long)(align_us / 1e3 + 0.5) << " msec" << std::endl;
    printf("Time elapsed (worker set-up before the bracket: thread start, device binding, output file creation): %.3f msec\n", setup_ms);
    int rcode = 0;
    for (auto &x : w) {
        int st = 0;
        waitpid(x.pid, &st, 0);
        if (!(WIFEXITED(st) && WEXITSTATUS(st) == 0)) rcode = 3;
        close(x.ctl); close(x.rep);
    }
    mark("workers_exited");
    long sorted_ms = -1;
    size_t sorted_lines = 0;
    if (const char *sp = getenv("DARWIN_SORTED_OUT")) {
        const auto ts = Clock::now();
        std::vector<std::string> lines;
        for (int g = 0; g < G; g++) {
            std::ifstream in("darwin." + std::to_string(g) + ".out");
            std::string ln;
            while (std::getline(in, ln)) lines.push_back(ln);
        }
        std::sort(lines.begin(), lines.end());
        lines.erase(std::unique(lines.begin(), lines.end()), lines.end());
        std::ofstream so(sp);
        for (auto &ln : lines) so << ln << '\n';
        so.close();
        sorted_lines = lines.size();
        sorted_ms = ms_since(ts);
        std::cout << "Time elapsed (sorted unique overlap file, " << sorted_lines << " lines): " << sorted_ms << " msec" << std::endl;
    }
    const double wall_s = std::chrono::duration<double>(Clock::now() - t_prog).count();
    std::cout << "Time elapsed (program start to here): " << (long)(wall_s * 1e3 + 0.5) << " msec" << std::endl;
    unsigned long long tiles = 0, cells = 0, cand = 0;
    double dev_ms = 0, sched_ms = 0;
    int batches = 0;
    for (auto &x : w) {
        tiles += (unsigned long long)json_num(x.done, "tiles"); cells += (unsigned long long)json_num(x.done, "cells");
        cand += (unsigned long long)json_num(x.done, "candidates");
        dev_ms = std::max(dev_ms, json_num(x.done, "gact_kernel_ms"));
        sched_ms = std::max(sched_ms, json_num(x.done, "gact_sched_ms"));
        batches = std::max(batches, (int)json_num(x.done, "chain_batches"));
        if (json_num(x.done, "error") != 0) rcode = 3;
    }
    {
        std::string tl = "DARWIN_B200_TIMELINE {";
        for (size_t i = 0; i < timeline.size(); i++) {
            char buf[160];
            snprintf(buf, sizeof(buf), "%s\"%s\": %.1f", i ? ", " : "", timeline[i].first.c_str(), timeline[i].second);
            tl += buf;
        }
        tl += "}";
        puts(tl.c_str());
    }
    printf("DARWIN_B200_SUMMARY {\"reads\": %lld, \"gpus\": %d, \"candidates\": %llu, \"tiles\": %llu, \"cells\": %llu, "
           "\"align_phase_ms\": %.3f, \"worker_setup_ms\": %.3f, \"gact_sched_ms\": %.1f, \"gact_kernel_ms\": %.1f, "
           "\"gpu_init_ms\": %.0f, \"seed_table_ms\": %.1f, \"teardown_ms\": %ld, \"chain_batches\": %d, \"wall_s\": %.3f, "
           "\"sorted_unique_lines\": %zu, \"worker_processes\": %d}\n",
           (long long)json_num(r0, "reads"), G, cand, tiles, cells, align_us / 1e3, setup_ms, sched_ms, dev_ms, init_ms, table_ms, 0L, batches,
           wall_s, sorted_lines, G);
    (void)same_file; (void)argc;
    return rcode;
}

int main(int argc, char **argv)
{
    if (argc < 4) {
        fprintf(stderr, "Usage: ./darwin <REFERENCE>.fasta <READS>.fasta CPU_THREADS [NUM_BLOCKS THREADS_PER_BLOCK]\n");
        return 1;
    }
    Params cfg;
    try {
        cfg = Params::from_file("params.cfg");
    } catch (const std::exception &e) {
        fprintf(stderr, "params.cfg: %s\n", e.what());
        return 1;
    }
    const std::string ref_path(argv[1]), reads_path(argv[2]);
    int num_threads = std::max(1, atoi(argv[3]));
    const bool same_file = (ref_path == reads_path);            // darwin.cpp:498-503
    // worker process of a multi-GPU run (see coordinate()): shard wk_index of wk_count, fd 3 = "go", fd 4 = reports
    int wk_index = -1, wk_count = 0;
    if (const char *wenv = getenv("DARWIN_WORKER")) {
        if (sscanf(wenv, "%d %d", &wk_index, &wk_count) != 2 || wk_index < 0 || wk_index >= wk_count) {
            fprintf(stderr, "darwin: bad DARWIN_WORKER\n");
            return 1;
        }
        unsetenv("DARWIN_WORKER");
        g_shard_fd = dup(1);
        if (!freopen("/dev/null", "w", stdout)) return 1;
        num_threads = std::max(1, num_threads / wk_count);
    }
    const bool is_worker = wk_index >= 0;
    // The CUDA driver enumerates every visible GPU when it starts (5 s on an 8-GPU box): a worker process, and a run that
    // uses one GPU, only need their own device -- narrow CUDA_VISIBLE_DEVICES before the first CUDA call.
    int device_base = 0;                                   // device index of shard 0 inside this process
    {
        const int want = getenv("DARWIN_GPUS") ? atoi(getenv("DARWIN_GPUS")) : 0;
        const int mine = is_worker ? wk_index : (want == 1 ? 0 : -1);
        if (mine >= 0) {
            std::string pick = std::to_string(mine);
            if (const char *vis = getenv("CUDA_VISIBLE_DEVICES")) {
                std::vector<std::string> ids;
                std::string cur;
                for (const char *c = vis;; c++) {
                    if (*c == ',' || *c == 0) { ids.push_back(cur); cur.clear(); if (*c == 0) break; }
                    else cur.push_back(*c);
                }
                pick = (size_t)mine < ids.size() ? ids[(size_t)mine] : std::string("none");
            }
            setenv("CUDA_VISIBLE_DEVICES", pick.c_str(), 1);
            device_base = -mine;                           // shard `mine` is device 0 of this process
        }
    }
    auto report = [&](const std::string &line) {               // worker -> coordinator
        const std::string l = line + "\n";
        ssize_t r = write(4, l.data(), l.size());
        (void)r;
    };
    if (!is_worker && getenv("DARWIN_GPUS") && atoi(getenv("DARWIN_GPUS")) > 1 &&
        !(getenv("DARWIN_MULTIPROC") && atoi(getenv("DARWIN_MULTIPROC")) == 0)) {
        printf("same_file: %d\n", same_file ? 1 : 0);
        return coordinate(argc, argv, atoi(getenv("DARWIN_GPUS")), cfg, num_threads, same_file);
    }
    printf("same_file: %d\n", same_file ? 1 : 0);

    // the device query initialises the CUDA driver (hundreds of ms): run it beside the FASTA loading
    const auto t_prog = Clock::now();
    // checkpoints of the whole run in ms since program start (printed as one DARWIN_B200_TIMELINE line)
    std::vector<std::pair<const char *, double>> timeline;
    std::mutex timeline_m;
    auto mark = [&](const char *what) {
        std::lock_guard<std::mutex> lk(timeline_m);
        timeline.emplace_back(what, us_since(t_prog) / 1e3);
    };
    int ndev = 0;
    long devq_ms = 0;
    std::thread devq([&] { ndev = gact_device_count(); devq_ms = ms_since(t_prog); });
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } devq_guard{devq};   // early error returns
    const int kernel_variant = getenv("DARWIN_KERNEL") ? atoi(getenv("DARWIN_KERNEL")) : 0;
    const bool use_chains = !(getenv("DARWIN_CHAINS") && atoi(getenv("DARWIN_CHAINS")) == 0);
    const bool dsoft_on_gpu = !(getenv("DARWIN_DSOFT") && std::string(getenv("DARWIN_DSOFT")) == "host");
    // seed-position table: built on every GPU by default (gact_seed_table_build); DARWIN_SEEDTABLE=host builds it
    // with the host threads and uploads it (the host D-SOFT path needs the host table in any case)
    const bool table_on_gpu = dsoft_on_gpu && !(getenv("DARWIN_SEEDTABLE") && std::string(getenv("DARWIN_SEEDTABLE")) == "host");
    printf("CPU threads: %d\n", num_threads);
    printf("Scores: match = %d, mismatch = %d, gap_open = %d, gap_extend = %d\n", cfg.match, cfg.mismatch, cfg.gap_open, cfg.gap_extend);
    printf("Minimizer window size: %d\n", (int)cfg.window_size);

    // ---- reference ------------------------------------------------------------------------
    std::cout << "\nLoading reference genome ...\n";
    auto t0 = Clock::now();
    FastaSet ref, reads;
    std::string err;
    if (!read_fasta(ref_path, &ref, &err)) { std::cerr << err << std::endl; return 1; }
    // concatenated reference, every sequence padded with 'N' to a multiple of bin_size (darwin.cpp:530-543)
    std::string ref_string;
    std::vector<uint32_t> chr_start_bin(ref.seqs.size());
    std::vector<int32_t> bin_to_chr;
    for (size_t i = 0; i < ref.seqs.size(); i++) {
        chr_start_bin[i] = (uint32_t)bin_to_chr.size();
        ref_string += ref.seqs[i];
        const size_t L = ref.seqs[i].size();
        for (size_t b = 0; b < L / cfg.bin_size; b++) bin_to_chr.push_back((int32_t)i);
        if (L % cfg.bin_size) {
            ref_string.append(cfg.bin_size - L % cfg.bin_size, 'N');
            bin_to_chr.push_back((int32_t)i);
        }
    }
    const uint32_t reference_length = (uint32_t)ref_string.size();
    std::cout << "Reference length: " << reference_length << ", " << ref.seqs.size() << " pieces" << std::endl;
    const long load_ref_ms = ms_since(t0);
    std::cout << "Time elapsed (loading reference genome): " << load_ref_ms << " msec" << std::endl;

    // ---- reads ----------------------------------------------------------------------------
    std::cout << "\nLoading reads ...\n";
    t0 = Clock::now();
    if (!read_fasta(reads_path, &reads, &err)) { std::cerr << err << std::endl; return 1; }
    const size_t num_reads = reads.seqs.size();
    std::vector<std::string> rev_reads(num_reads);
    for (size_t i = 0; i < num_reads; i++) {
        char bad = 0;
        if (!reverse_complement(reads.seqs[i], &rev_reads[i], &bad)) {
            std::cerr << "Bad Nt char: " << bad << std::endl;       // darwin.cpp:117-119
            return 1;
        }
    }
    std::cout << "Number of reads: " << num_reads << std::endl;
    const long load_reads_ms = ms_since(t0);
    std::cout << "Time elapsed (loading reads): " << load_reads_ms << " msec" << std::endl;
    mark("fasta_loaded");

    devq.join();
    mark("cuda_driver_up");
    if (ndev <= 0) {
        fprintf(stderr, "darwin: no CUDA device available (the GACT path has no CPU fallback)\n");
        return 2;
    }
    int want_gpus = ndev;
    if (const char *e = getenv("DARWIN_GPUS")) want_gpus = std::max(1, std::min(ndev, atoi(e)));
    if (is_worker) {
        want_gpus = wk_count;                              // this process sees its own device only
    }
    printf("Using GPU: %d device(s); CUDA driver initialised after %ld msec\n", want_gpus, devq_ms);

    // ---- shards: contiguous read ranges, one per GPU (darwin.cpp:619-629 rule) ----------------
    const int G = (int)std::max<size_t>(1, std::min<size_t>((size_t)want_gpus, std::max<size_t>(num_reads, 1)));
    const size_t per = (num_reads + G - 1) / std::max(G, 1);
    std::vector<Shard> shards;
    for (int g = 0; g < G; g++) {
        Shard s;
        s.tid = g; s.device = g + device_base;
        s.first_read = std::min(num_reads, per * g);
        s.last_read = std::min(num_reads, per * (g + 1));
        s.dsoft_threads = is_worker ? num_threads : std::max(1, num_threads / G);
        if (is_worker && g != wk_index) continue;          // a worker process owns exactly one shard
        if (s.first_read < s.last_read || g == 0) shards.push_back(s);
    }
    if (is_worker && shards.empty()) {
        // more GPUs than reads: nothing to do for this worker, but the coordinator still waits for its two reports
        report("READY {\"reads\": " + std::to_string(num_reads) + ", \"idle\": 1}");
        char go = 0;
        ssize_t r = read(3, &go, 1);
        (void)r;
        report("DONE {\"tiles\": 0, \"cells\": 0, \"candidates\": 0, \"error\": 0}");
        return 0;
    }

    // ---- engines: one per shard/GPU, created and loaded while the seed table is being built -------
    const gact_params gp{cfg.match, cfg.mismatch, cfg.gap_open, cfg.gap_extend, cfg.tile_size, cfg.tile_overlap,
                         cfg.first_tile_score_threshold};
    auto t_gpu = Clock::now();
    std::vector<std::thread> init_threads;
    for (auto &sh : shards) {
        Shard *shp = &sh;
        init_threads.emplace_back([&, shp] {
            Shard &S = *shp;
            const auto t_init = Clock::now();
            int rc = gact_engine_create(&S.eng, S.device, &gp, 1 << 17, nullptr);
            if (rc) { S.error = std::string("gact_engine_create: ") + gact_last_error(nullptr); return; }
            if (kernel_variant) gact_engine_set_kernel(S.eng, kernel_variant);
            std::vector<const char *> ptrs;
            std::vector<int64_t> lens;
            auto upload = [&](int set, const std::vector<std::string> &v, size_t a, size_t b) {
                ptrs.clear(); lens.clear();
                for (size_t i = a; i < b; i++) { ptrs.push_back(v[i].data()); lens.push_back((int64_t)v[i].size()); }
                if (gact_engine_upload(S.eng, set, (int64_t)ptrs.size(), ptrs.data(), lens.data()) && S.error.empty())
                    S.error = std::string("gact_engine_upload: ") + gact_last_error(S.eng);
            };
            upload(GACT_SET_REF, ref.seqs, 0, ref.seqs.size());
            upload(GACT_SET_READS, reads.seqs, S.first_read, S.last_read);
            upload(GACT_SET_READS_RC, rev_reads, S.first_read, S.last_read);
            bool slow_reads = false;                      // reads with bytes other than ACGT take the tile-by-tile path
            S.read_exc.assign(S.last_read - S.first_read, 0);
            for (size_t k = 0; k < S.last_read - S.first_read; k++) {
                S.read_exc[k] = gact_engine_seq_has_exceptions(S.eng, GACT_SET_READS, (int64_t)k) == 1;
                slow_reads |= S.read_exc[k] != 0;
            }
            if (use_chains && gact_engine_extend_supported(S.eng)) gact_engine_extend_reserve(S.eng, (int)(8 * (S.last_read - S.first_read) + 1024));
            if (slow_reads || !(use_chains && gact_engine_extend_supported(S.eng)))
                gact_engine_reserve_tiles(S.eng);         // tile-by-tile host scheduler: its batch slots belong to initialisation
            S.init_ms = ms_since(t_init);
        });
    }

    // ---- seed table -----------------------------------------------------------------------
    std::cout << "\nConstructing seed position table ...\n";
    t0 = Clock::now();
    SeedTable *table = nullptr;
    if (!table_on_gpu) {
        try {
            table = new SeedTable(ref_string.data(), reference_length, cfg.seed_size, (uint32_t)cfg.seed_occurence_multiple,
                                  cfg.bin_size, cfg.window_size, num_threads);
        } catch (const std::exception &e) {
            fprintf(stderr, "seed table: %s\n", e.what());
            return 1;
        }
        std::cout << "Time elapsed (seed position table construction): " << ms_since(t0) << " msec" << std::endl;
    }

    for (auto &th : init_threads) th.join();
    mark("engines_created_sequences_uploaded");
    std::cout << "Time elapsed (GPU init" << (table_on_gpu ? "" : ", overlapped with the seed table") << "): " << ms_since(t_gpu)
              << " msec" << std::endl;
    for (auto &sh : shards) std::cout << "GPU " << sh.device << " init alone (context, engine, sequence upload): " << sh.init_ms << " msec" << std::endl;
    for (auto &sh : shards)
        if (!sh.error.empty()) { fprintf(stderr, "shard %d: %s\n", sh.tid, sh.error.c_str()); return 3; }

    if (dsoft_on_gpu) {
        // every GPU gets the seed-position table once (index: 4^k + 1 words, positions: one word per minimizer):
        // built there from the reference string, or uploaded from the host builder
        auto t_up = Clock::now();
        std::vector<std::thread> up;
        for (auto &sh : shards) {
            Shard *shp = &sh;
            up.emplace_back([&, shp] {
                int rc;
                if (table_on_gpu) {
                    rc = gact_seed_table_build(&shp->seed_table, shp->eng, ref_string.data(), reference_length, cfg.seed_size,
                                               (uint32_t)cfg.seed_occurence_multiple, cfg.bin_size, cfg.window_size);
                    if (rc) { shp->error = std::string("gact_seed_table_build: ") + gact_last_error(shp->eng); return; }
                    gact_seed_table_info(shp->seed_table, nullptr, nullptr, nullptr, &shp->table_kernel_ms);
                    rc = gact_dsoft_create_from_table(&shp->dsoft, shp->eng, shp->seed_table, cfg.num_seeds, cfg.threshold,
                                                      cfg.max_candidates);
                } else {
                    rc = gact_dsoft_create(&shp->dsoft, shp->eng, table->index_table(), table->index_entries(), table->pos_table(),
                                           table->num_minimizers(), cfg.seed_size, (int)cfg.window_size, cfg.bin_size,
                                           table->kmer_max_occurence(), cfg.num_seeds, cfg.threshold, cfg.max_candidates);
                }
                if (rc) shp->error = std::string("gact_dsoft_create: ") + gact_last_error(shp->eng);
                // buffers for the timed phase: two queries per read, a few candidates per read
                const size_t nrs = shp->last_read - shp->first_read;
                if (!rc) gact_dsoft_reserve(shp->dsoft, (int)(2 * nrs), (int64_t)std::max<size_t>(1024, 8 * nrs));
                // one throw-away query and one throw-away extension: the kernels' code is loaded onto the device at
                // their first launch, which belongs to initialisation like the rest of GPU_init (darwin.cpp:611)
                if (!rc && nrs > 0 && !ref.seqs.empty()) {
                    const int32_t qs = GACT_SET_READS;
                    const int64_t qi = 0;
                    gact_dsoft_cand tmp[64];
                    int64_t n_tmp = 0;
                    gact_dsoft_run(shp->dsoft, 1, &qs, &qi, tmp, 64, &n_tmp);
                    if (use_chains && gact_engine_extend_supported(shp->eng)) {
                        gact_call wc;
                        memset(&wc, 0, sizeof(wc));
                        wc.query_set = GACT_SET_READS;
                        wc.ref_pos = (int32_t)std::min<size_t>(ref.seqs[0].size(), 64);
                        wc.query_pos = (int32_t)std::min<size_t>(reads.seqs[shp->first_read].size(), 64);
                        gact_alignment wa;
                        gact_engine_extend(shp->eng, 1, &wc, &wa);
                    }
                    gact_engine_reset_stats(shp->eng);
                }
            });
        }
        for (auto &th : up) th.join();
        mark("seed_tables_and_filters_ready");
        if (table_on_gpu) {
            std::cout << "Time elapsed (seed position table construction): " << ms_since(t0) << " msec" << std::endl;
            for (auto &sh : shards)
                std::cout << "GPU " << sh.device << " seed table build (H2D + kernels): " << sh.table_kernel_ms << " msec" << std::endl;
        } else {
            std::cout << "Time elapsed (seed table upload to GPU): " << ms_since(t_up) << " msec" << std::endl;
        }
        for (auto &sh : shards)
            if (!sh.error.empty()) { fprintf(stderr, "shard %d: %s\n", sh.tid, sh.error.c_str()); return 3; }
    }

    std::vector<SeqView> ref_views(ref.seqs.size());
    for (size_t i = 0; i < ref.seqs.size(); i++) ref_views[i] = SeqView{ref.seqs[i].data(), (int64_t)ref.seqs[i].size()};

    std::cout << "\nFinding candidate bin locations for each read: " << std::endl;

    // Pipelined shard (GPU D-SOFT + on-device chains).  The D-SOFT filter runs once for the whole shard (half a millisecond of
    // kernel time for 50 MB of reads); the candidates then go through gact_engine_extend_submit / _wait in batches of
    // consecutive reads, up to GACT_MAX_INFLIGHT batches on their own streams, and a writer thread formats and writes
    // batch k while the chains of the following batches run.
    size_t batch_reads_env = 0;
    bool batch_reads_set = false;
    if (const char *b = getenv("DARWIN_BATCH_READS")) { batch_reads_env = (size_t)std::max(0, atoi(b)); batch_reads_set = true; }
    auto run_shard_pipelined = [&](Shard &sh, std::ofstream &fout) {
        try {
            const size_t nr = sh.last_read - sh.first_read;
            if (!fout.is_open()) { sh.error = "ERROR cannot open output file"; return; }
            const auto td = Clock::now();
            gact_engine *eng = sh.eng;
            // DARWIN_TRACE=1: microsecond checkpoints of this shard's align phase (where the host time goes)
            const bool trace = getenv("DARWIN_TRACE") != nullptr;
            std::string trace_line;
            auto tmark = [&](const char *what) {
                if (trace) { char b[96]; snprintf(b, sizeof(b), " %s %.0f", what, us_since(td)); trace_line += b; }
            };

            // ---- D-SOFT: queries = (reads, k), (reverse-complemented reads, k) for every read of the shard ----
            std::vector<int32_t> qsets(2 * nr);
            std::vector<int64_t> qidx(2 * nr);
            for (size_t k = 0; k < nr; k++) {
                qsets[2 * k] = GACT_SET_READS; qsets[2 * k + 1] = GACT_SET_READS_RC;
                qidx[2 * k] = qidx[2 * k + 1] = (int64_t)k;
            }
            std::vector<gact_dsoft_cand> cands(std::max<size_t>(1024, 8 * nr));
            int64_t n_cands = 0;
            int rc = gact_dsoft_run(sh.dsoft, (int)(2 * nr), qsets.data(), qidx.data(), cands.data(), (int64_t)cands.size(), &n_cands);
            if (rc == GACT_ERR_NOMEM && n_cands > (int64_t)cands.size()) {
                cands.resize((size_t)n_cands);
                rc = gact_dsoft_run(sh.dsoft, (int)(2 * nr), qsets.data(), qidx.data(), cands.data(), (int64_t)cands.size(), &n_cands);
            }
            if (rc) { sh.error = std::string("gact_dsoft_run: ") + gact_last_error(eng); return; }
            sh.dsoft_wait_us = us_since(td);
            tmark("dsoft_done");

            // ---- batches of consecutive reads ----
            // one batch unless the shard has very many candidates (de-novo self-alignments: 200 k for config 1): the chains of
            // one batch pack best (longest first over the whole shard); several batches only pay where the writer's work and
            // the result copies of batch k are worth hiding behind the chains of batch k+1
            size_t per_batch = nr;
            if (batch_reads_set) per_batch = batch_reads_env ? batch_reads_env : nr;
            else if (n_cands > 49152) { const size_t nb = ((size_t)n_cands + 32767) / 32768; per_batch = (nr + nb - 1) / nb; }
            per_batch = std::max<size_t>(per_batch, 1);
            const size_t B = std::max<size_t>(1, (nr + per_batch - 1) / per_batch);
            sh.batches = (int)B;
            // calls: every candidate of the batch in output order; gc: those that go through gact_engine_extend (gc_pos: their
            // index in calls).  A candidate whose read holds a byte other than ACGT goes through the tile-by-tile host
            // scheduler instead (raw-byte kernels), the reference's raw byte comparison needs that (align.cpp:134).
            struct Batch { std::vector<GactCall> calls; std::vector<gact_call> gc; std::vector<size_t> gc_pos;
                           std::vector<gact_alignment> ga, ga_dev; };
            std::vector<Batch> bt(B);
            std::vector<GactCall> slow_calls;
            std::vector<std::pair<size_t, size_t>> slow_at;           // (batch, index in calls)
            // candidates arrive grouped by query (= read, strand) in emission order, i.e. in the reference CPU build's
            // order per read: forward candidates, then reverse (darwin.cpp:209-288)
            {
                std::vector<size_t> per((size_t)B, 0);
                for (int64_t i = 0; i < n_cands; i++) per[std::min(B - 1, (size_t)(cands[(size_t)i].query >> 1) / per_batch)]++;
                for (size_t b = 0; b < B; b++) { bt[b].calls.reserve(per[b]); bt[b].gc.reserve(per[b]); bt[b].gc_pos.reserve(per[b]); }
            }
            for (int64_t i = 0; i < n_cands; i++) {
                const gact_dsoft_cand &c = cands[(size_t)i];
                const bool comp = (c.query & 1) != 0;
                const size_t local = (size_t)(c.query >> 1);                              // read index inside the shard
                int ref_pos = (int)c.hit;
                const int chr = bin_to_chr[(uint32_t)ref_pos / cfg.bin_size];
                ref_pos -= (int)(chr_start_bin[chr] * cfg.bin_size);
                if (ref_pos > (long long)ref.seqs[chr].size()) ref_pos = (int)ref.seqs[chr].size();       // darwin.cpp:222-224
                (comp ? sh.cand_rev : sh.cand_fwd)++;
                const size_t bi = std::min(B - 1, local / per_batch);
                Batch &x = bt[bi];
                x.calls.push_back(GactCall{chr, (int32_t)local, ref_pos, (int)c.offset, (uint8_t)(comp ? 1 : 0)});
                if (sh.read_exc[local]) {
                    slow_calls.push_back(x.calls.back());
                    slow_at.emplace_back(bi, x.calls.size() - 1);
                    continue;
                }
                gact_call g;
                memset(&g, 0, sizeof(g));
                g.ref_seq = chr; g.query_seq = (int32_t)local; g.ref_pos = ref_pos; g.query_pos = (int)c.offset;
                g.query_set = comp ? GACT_SET_READS_RC : GACT_SET_READS;
                x.gc.push_back(g);
                x.gc_pos.push_back(x.calls.size() - 1);
            }
            for (Batch &x : bt) { x.ga.resize(x.calls.size()); x.ga_dev.resize(x.gc.size()); }
            tmark("calls_built");
            sh.text.assign(B, std::string());
            sh.dsoft_ms = ms_since(td);
            shard_print("num_candidates: " + std::to_string(sh.cand_fwd) + " " + std::to_string(sh.cand_rev));
            shard_print("Time finding seeds: " + std::to_string(sh.dsoft_ms) + " msec");
            const auto tg = Clock::now();
            if (!slow_calls.empty()) {
                // reads with exceptions: tile by tile through the host scheduler, before the chain batches are in flight
                std::vector<SeqView> rd(nr), rc_views(nr);
                for (size_t k = 0; k < nr; k++) {
                    rd[k] = SeqView{reads.seqs[sh.first_read + k].data(), (int64_t)reads.seqs[sh.first_read + k].size()};
                    rc_views[k] = SeqView{rev_reads[sh.first_read + k].data(), (int64_t)rev_reads[sh.first_read + k].size()};
                }
                std::vector<GactAlignment> slow_aln;
                GactScheduler sched(eng, gp, ref_views, rd, rc_views, sh.dsoft_threads);
                sched.run(slow_calls, slow_aln, &sh.stats);
                for (size_t k = 0; k < slow_calls.size(); k++) {
                    const GactAlignment &a = slow_aln[k];
                    gact_alignment g;
                    memset(&g, 0, sizeof(g));
                    g.ab = a.ab; g.ae = a.ae; g.bb = a.bb; g.be = a.be; g.score = a.score; g.first_tile_score = a.first_tile_score;
                    g.n_tiles = a.n_tiles; g.n_cells = a.n_cells;
                    bt[slow_at[k].first].ga[slow_at[k].second] = g;
                }
            }

            // writer thread: batches in order, same line order as the sequential loop of the reference
            std::mutex wm;
            std::condition_variable wcv;
            size_t ready = 0;                         // batches whose alignments are complete
            bool stop_writer = false;
            std::thread writer([&] {
                for (size_t b = 0; b < B; b++) {
                    {
                        std::unique_lock<std::mutex> lk(wm);
                        wcv.wait(lk, [&] { return ready > b || stop_writer; });
                        if (ready <= b) return;       // aborted
                    }
                    const auto tw = Clock::now();
                    const Batch &x = bt[b];
                    // formatted in parallel chunks (same line order as the sequential loop), one write per batch
                    const int parts = std::max(1, std::min(sh.dsoft_threads, (int)(x.calls.size() / 1024) + 1));
                    std::vector<std::string> piece((size_t)parts);
                    auto fmt = [&](int p) {
                        const size_t a0 = x.calls.size() * (size_t)p / parts, a1 = x.calls.size() * (size_t)(p + 1) / parts;
                        std::string &o = piece[(size_t)p];
                        o.reserve((a1 - a0) * 112);
                        for (size_t k = a0; k < a1; k++) {
                            const GactCall &c = x.calls[k];
                            const size_t read_id = sh.first_read + (size_t)c.query_id;
                            const gact_alignment &a = x.ga[k];
                            if (!(same_file && (size_t)c.ref_id == read_id) && a.score > 0)        // gact.cpp:213
                                o += format_overlap(ref.names[c.ref_id], reads.names[read_id],
                                                    GactAlignment{a.ab, a.ae, a.bb, a.be, a.score, a.first_tile_score, a.n_tiles, a.n_cells},
                                                    c.complement != 0);
                        }
                    };
                    std::vector<std::thread> helpers;
                    for (int p = 1; p < parts; p++) helpers.emplace_back(fmt, p);
                    fmt(0);
                    for (auto &h : helpers) h.join();
                    std::string &out = sh.text[b];
                    size_t total = 0;
                    for (auto &o : piece) total += o.size();
                    out.reserve(total);
                    for (auto &o : piece) out += o;
                    fout.write(out.data(), (std::streamsize)out.size());
                    sh.write_us += us_since(tw);
                }
                fout.close();
            });
            auto finish_writer = [&](bool abort) {
                { std::lock_guard<std::mutex> lk(wm); if (abort) stop_writer = true; }
                wcv.notify_all();
                if (writer.joinable()) writer.join();
            };
            auto extend_collect = [&](size_t b) -> int {
                const auto tw = Clock::now();
                const int r = gact_engine_extend_wait(eng, bt[b].ga_dev.data());
                sh.extend_wait_us += us_since(tw);
                if (r) return r;
                for (size_t k = 0; k < bt[b].ga_dev.size(); k++) {
                    const gact_alignment &a = bt[b].ga_dev[k];
                    bt[b].ga[bt[b].gc_pos[k]] = a;
                    sh.stats.tiles += (uint64_t)a.n_tiles; sh.stats.cells += (uint64_t)a.n_cells;
                }
                { std::lock_guard<std::mutex> lk(wm); ready = b + 1; }
                wcv.notify_all();
                return GACT_OK;
            };
            size_t collected = 0;
            rc = GACT_OK;
            for (size_t b = 0; b < B && !rc; b++) {
                while (!rc && b - collected >= (size_t)GACT_MAX_INFLIGHT) rc = extend_collect(collected++);
                if (!rc) rc = gact_engine_extend_submit(eng, (int)bt[b].gc.size(), bt[b].gc.data());
                tmark("submitted");
            }
            while (!rc && collected < B) { rc = extend_collect(collected++); tmark("collected"); }
            if (rc) { finish_writer(true); sh.error = std::string("gact_engine_extend: ") + gact_last_error(eng); return; }
            finish_writer(false);
            tmark("written");
            if (trace) shard_print("TRACE shard " + std::to_string(sh.tid) + " (us since the shard started):" + trace_line);
            gact_stats es;
            gact_engine_stats(eng, &es);
            sh.stats.device_ms = es.kernel_ms;
            sh.stats.rounds = B;
            sh.gact_ms = ms_since(tg);
            sh.stats.wall_ms = (double)sh.gact_ms;
            shard_print("Time GACT calling: " + std::to_string(sh.gact_ms) + " msec (" + std::to_string(B) + " batch(es) of chains; host blocked on chains "
                        + std::to_string((long)(sh.extend_wait_us / 1e3 + 0.5)) + " msec; writer thread busy "
                        + std::to_string((long)(sh.write_us / 1e3 + 0.5)) + " msec)");
        } catch (const std::exception &e) {
            sh.error = e.what();
        }
    };

    auto run_shard = [&](Shard &sh, std::ofstream &fout) {
        try {
            const size_t nr = sh.last_read - sh.first_read;
            if (!fout.is_open()) { sh.error = "ERROR cannot open output file"; return; }

            // D-SOFT for every read of the shard, both strands (darwin.cpp:209-288)
            auto td = Clock::now();
            std::vector<std::vector<uint64_t>> cand_f(nr), cand_r(nr);
            if (sh.dsoft) {
                // device-side filter: queries = (reads, k), (reverse-complemented reads, k) for every read of the shard
                std::vector<int32_t> qsets(2 * nr);
                std::vector<int64_t> qidx(2 * nr);
                for (size_t k = 0; k < nr; k++) {
                    qsets[2 * k] = GACT_SET_READS; qsets[2 * k + 1] = GACT_SET_READS_RC;
                    qidx[2 * k] = qidx[2 * k + 1] = (int64_t)k;
                }
                std::vector<gact_dsoft_cand> cands(std::max<size_t>(1024, 8 * nr));
                int64_t n_out = 0;
                int rc = gact_dsoft_run(sh.dsoft, (int)(2 * nr), qsets.data(), qidx.data(), cands.data(), (int64_t)cands.size(), &n_out);
                if (rc == GACT_ERR_NOMEM && n_out > (int64_t)cands.size()) {
                    cands.resize((size_t)n_out);
                    rc = gact_dsoft_run(sh.dsoft, (int)(2 * nr), qsets.data(), qidx.data(), cands.data(), (int64_t)cands.size(), &n_out);
                }
                if (rc) { sh.error = std::string("gact_dsoft_run: ") + gact_last_error(sh.eng); return; }
                for (int64_t x = 0; x < n_out; x++) {
                    const gact_dsoft_cand &c = cands[(size_t)x];
                    auto &dst = (c.query & 1) ? cand_r[(size_t)(c.query >> 1)] : cand_f[(size_t)(c.query >> 1)];
                    dst.push_back(((uint64_t)c.hit << 32) | c.offset);
                }
            } else {
                std::vector<std::thread> th;
                std::atomic<size_t> next(0);
                for (int t = 0; t < sh.dsoft_threads; t++) th.emplace_back([&] {
                    SeedTable::Scratch scratch(*table, cfg.num_nz_bins);
                    for (;;) {
                        const size_t k = next.fetch_add(1);
                        if (k >= nr) break;
                        const size_t r = sh.first_read + k;
                        table->dsoft(reads.seqs[r].data(), (uint32_t)reads.seqs[r].size(), cfg.num_seeds, cfg.threshold,
                                     cfg.max_candidates, scratch, cand_f[k]);
                        table->dsoft(rev_reads[r].data(), (uint32_t)rev_reads[r].size(), cfg.num_seeds, cfg.threshold,
                                     cfg.max_candidates, scratch, cand_r[k]);
                    }
                });
                for (auto &x : th) x.join();
            }
            // candidates -> calls, in the reference CPU build's order: per read forward then reverse
            std::vector<GactCall> calls;
            auto add_calls = [&](const std::vector<uint64_t> &cands, size_t read_id, bool comp) {
                for (uint64_t c : cands) {
                    int ref_pos = (int)(c >> 32);
                    const int chr = bin_to_chr[(uint32_t)ref_pos / cfg.bin_size];
                    ref_pos -= (int)(chr_start_bin[chr] * cfg.bin_size);
                    const int query_pos = (int)(c & 0xffffffffu);
                    if (ref_pos > (long long)ref.seqs[chr].size()) ref_pos = (int)ref.seqs[chr].size();   // darwin.cpp:222-224
                    calls.push_back(GactCall{chr, (int32_t)read_id, ref_pos, query_pos, (uint8_t)(comp ? 1 : 0)});
                }
            };
            for (size_t k = 0; k < nr; k++) {
                sh.cand_fwd += cand_f[k].size();
                sh.cand_rev += cand_r[k].size();
                add_calls(cand_f[k], sh.first_read + k, false);
                add_calls(cand_r[k], sh.first_read + k, true);
            }
            sh.dsoft_ms = ms_since(td);
            {
                std::lock_guard<std::mutex> lk(io_lock);
                printf("num_candidates: %llu %llu\n", (unsigned long long)sh.cand_fwd, (unsigned long long)sh.cand_rev);
                std::cout << "Time finding seeds: " << sh.dsoft_ms << " msec" << std::endl;
            }

            auto tg = Clock::now();
            gact_engine *eng = sh.eng;
            std::vector<SeqView> rd(nr), rc_views(nr);
            for (size_t k = 0; k < nr; k++) {
                rd[k] = SeqView{reads.seqs[sh.first_read + k].data(), (int64_t)reads.seqs[sh.first_read + k].size()};
                rc_views[k] = SeqView{rev_reads[sh.first_read + k].data(), (int64_t)rev_reads[sh.first_read + k].size()};
            }
            // query ids inside the engine are shard-local
            for (auto &c : calls) c.query_id -= (int32_t)sh.first_read;
            std::vector<GactAlignment> aln;
            if (use_chains && gact_engine_extend_supported(eng)) {
                // whole GACT() per candidate on the device; longest reads first (their chains are the longest)
                std::vector<uint32_t> ord(calls.size());
                for (size_t k = 0; k < ord.size(); k++) ord[k] = (uint32_t)k;
                std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) {
                    return rd[(size_t)calls[a].query_id].len > rd[(size_t)calls[b].query_id].len;
                });
                std::vector<gact_call> gc(calls.size());
                for (size_t k = 0; k < ord.size(); k++) {
                    const GactCall &c = calls[ord[k]];
                    gact_call g;
                    memset(&g, 0, sizeof(g));
                    g.ref_seq = c.ref_id; g.query_seq = c.query_id; g.ref_pos = c.ref_pos; g.query_pos = c.query_pos;
                    g.query_set = c.complement ? GACT_SET_READS_RC : GACT_SET_READS;
                    gc[k] = g;
                }
                std::vector<gact_alignment> ga(calls.size());
                auto tc = Clock::now();
                int rc2 = gact_engine_extend(eng, (int)gc.size(), gc.data(), ga.data());
                if (rc2) { sh.error = std::string("gact_engine_extend: ") + gact_last_error(eng); return; }
                sh.stats.wall_ms += std::chrono::duration<double, std::milli>(Clock::now() - tc).count();
                aln.resize(calls.size());
                for (size_t k = 0; k < ord.size(); k++) {
                    const gact_alignment &a = ga[k];
                    aln[ord[k]] = GactAlignment{a.ab, a.ae, a.bb, a.be, a.score, a.first_tile_score, a.n_tiles, a.n_cells};
                    sh.stats.tiles += (uint64_t)a.n_tiles;
                    sh.stats.cells += (uint64_t)a.n_cells;
                }
                sh.stats.rounds += 1;
            } else {
                GactScheduler sched(eng, gp, ref_views, rd, rc_views, sh.dsoft_threads);
                sched.run(calls, aln, &sh.stats);
            }
            gact_stats es;
            gact_engine_stats(eng, &es);
            sh.stats.device_ms = es.kernel_ms;
            const long t_sched = ms_since(tg);
            // the engine is torn down after the output is written (outside the per-shard critical path)

            // format in parallel chunks (same line order as the sequential loop), one write per chunk
            {
                const int parts = std::max(1, std::min(sh.dsoft_threads, (int)(calls.size() / 512) + 1));
                std::vector<std::string> text((size_t)parts);
                std::vector<std::thread> fmt;
                for (int p = 0; p < parts; p++) fmt.emplace_back([&, p] {
                    const size_t a = calls.size() * (size_t)p / parts, b = calls.size() * (size_t)(p + 1) / parts;
                    std::string &out = text[(size_t)p];
                    out.reserve((b - a) * 112);
                    for (size_t k = a; k < b; k++) {
                        const GactCall &c = calls[k];
                        const size_t read_id = sh.first_read + (size_t)c.query_id;
                        if (!(same_file && (size_t)c.ref_id == read_id) && aln[k].score > 0)        // gact.cpp:213
                            out += format_overlap(ref.names[c.ref_id], reads.names[read_id], aln[k], c.complement != 0);
                    }
                });
                for (auto &t : fmt) t.join();
                for (auto &t : text) fout.write(t.data(), (std::streamsize)t.size());
                sh.text.swap(text);
            }
            fout.close();
            sh.gact_ms = ms_since(tg);
            const long t_out = sh.gact_ms - t_sched;
            {
                std::lock_guard<std::mutex> lk(io_lock);
                std::cout << "Time GACT scheduler: " << t_sched << " msec, writing overlaps: " << t_out << " msec" << std::endl;
            }
            {
                std::lock_guard<std::mutex> lk(io_lock);
                std::cout << "Time GACT calling: " << sh.gact_ms << " msec" << std::endl;
            }
        } catch (const std::exception &e) {
            sh.error = e.what();
        }
    };

    // One worker per shard, as darwin.cpp:619-629.  The workers are started and touch their device once before the
    // timed bracket opens: a host thread's first CUDA call binds it to the device context, which costs milliseconds
    // per thread in a multi-GPU process and belongs to initialisation.
    std::vector<std::thread> workers;
    std::mutex gate_m;
    std::condition_variable gate_cv;
    size_t parked = 0, finished = 0;
    bool go = false;
    for (auto &sh : shards) {
        Shard *shp = &sh;
        const auto t_spawn = Clock::now();
        workers.emplace_back([&, shp, t_spawn] {
            gact_engine_sync(shp->eng);
            // output file of this worker (darwin.cpp:203-204), created before the bracket: the first file creation of the
            // process was measured at 12-40 ms on the test boxes' file system, next to a 10 ms alignment phase
            std::ofstream fout("darwin." + std::to_string(shp->tid) + ".out");
            shp->setup_us = us_since(t_spawn);
            {
                std::unique_lock<std::mutex> lk(gate_m);
                parked++;
                gate_cv.notify_all();
                gate_cv.wait(lk, [&] { return go; });
            }
            if (shp->dsoft && use_chains && gact_engine_extend_supported(shp->eng)) run_shard_pipelined(*shp, fout);
            else run_shard(*shp, fout);
            {
                std::lock_guard<std::mutex> lk(gate_m);
                finished++;
            }
            gate_cv.notify_all();
        });
    }
    {
        std::unique_lock<std::mutex> lk(gate_m);
        gate_cv.wait(lk, [&] { return parked == shards.size(); });
        if (is_worker) {
            // tell the coordinator that this GPU is ready, then wait for its "go": the bracket is the coordinator's
            char buf[512];
            snprintf(buf, sizeof(buf), "READY {\"reads\": %zu, \"reference_length\": %u, \"ref_pieces\": %zu, \"load_ref_ms\": %ld, "
                     "\"load_reads_ms\": %ld, \"cuda_up_ms\": %ld, \"gpu_init_ms\": %ld, \"seed_table_phase_ms\": %ld, \"seed_table_ms\": %.1f, "
                     "\"worker_setup_ms\": %.3f, \"ready_ms\": %.1f}", num_reads, reference_length, ref.seqs.size(), load_ref_ms, load_reads_ms,
                     devq_ms, shards[0].init_ms, ms_since(t_gpu), shards[0].table_kernel_ms, shards[0].setup_us / 1e3, us_since(t_prog) / 1e3);
            report(buf);
            char gobyte = 0;
            if (read(3, &gobyte, 1) != 1) _exit(4);          // the coordinator is gone
        }
        mark("workers_parked_bracket_opens");
        t0 = Clock::now();
        go = true;
        gate_cv.notify_all();
    }
    std::cout << workers.size() << " threads created\n";
    std::cout << "Synchronizing all threads...\n";
    long align_ms = 0;
    double align_us = 0;
    {
        // the bracket closes when every shard has written its output; the threads' exit (the CUDA runtime's
        // per-thread teardown) is joined afterwards, next to GPU_close
        std::unique_lock<std::mutex> lk(gate_m);
        gate_cv.wait(lk, [&] { return finished == shards.size(); });
        align_us = us_since(t0);
        align_ms = (long)(align_us / 1e3 + 0.5);
    }
    mark("bracket_closes");
    if (is_worker) {
        const Shard &sh = shards[0];
        if (!sh.error.empty()) fprintf(stderr, "shard %d: %s\n", sh.tid, sh.error.c_str());
        char buf[512];
        snprintf(buf, sizeof(buf), "DONE {\"tiles\": %llu, \"cells\": %llu, \"candidates\": %llu, \"align_phase_ms\": %.3f, "
                 "\"gact_sched_ms\": %.1f, \"gact_kernel_ms\": %.1f, \"chain_batches\": %d, \"error\": %d}",
                 (unsigned long long)sh.stats.tiles, (unsigned long long)sh.stats.cells, (unsigned long long)(sh.cand_fwd + sh.cand_rev),
                 align_us / 1e3, sh.stats.wall_ms, sh.stats.device_ms, sh.batches, sh.error.empty() ? 0 : 1);
        report(buf);
        // the output file is closed and reported; the driver releases this process's context when it exits, which is
        // faster than destroying engine, filter and seed table one by one (GPU_close comes after the reference's bracket too)
        fflush(nullptr);
        _exit(sh.error.empty() ? 0 : 3);
    }
    const auto t_join = Clock::now();
    for (auto &w : workers) w.join();
    const long join_ms = ms_since(t_join);
    mark("workers_joined");
    std::cout << "Time elapsed (seed table querying + aligning): " << align_ms << " msec" << std::endl;
    std::cout << "Time elapsed (worker thread exit): " << join_ms << " msec" << std::endl;
    // The reference starts its worker threads and opens the per-thread output files inside its bracket
    // (darwin.cpp:174-175, 615-639); here both happen before the bracket opens.  Their cost is reported so that the
    // like-for-like figure can be formed: bracket + slowest worker set-up.
    double setup_us = 0;
    for (auto &sh : shards) setup_us = std::max(setup_us, sh.setup_us);
    printf("Time elapsed (worker set-up before the bracket: thread start, device binding, output file creation): %.3f msec\n", setup_us / 1e3);
    // GPU_close comes after the timed phase in the reference as well (darwin.cpp:634-642)
    const auto t_down = Clock::now();
    for (auto &sh : shards) {
        if (sh.dsoft) { gact_dsoft_destroy(sh.dsoft); sh.dsoft = nullptr; }
        if (sh.seed_table) { gact_seed_table_destroy(sh.seed_table); sh.seed_table = nullptr; }
        if (sh.eng) { gact_engine_destroy(sh.eng); sh.eng = nullptr; }
    }
    const long down_ms = ms_since(t_down);
    mark("gpu_teardown_done");
    std::cout << "Time elapsed (GPU teardown): " << down_ms << " msec" << std::endl;

    // f4: the sorted, duplicate-free overlap list in one file (README:32 builds it with `cat darwin.*.out | sort | uniq`;
    // byte order = `LC_ALL=C sort`)
    long sorted_ms = -1;
    size_t sorted_lines = 0;
    if (const char *sp = getenv("DARWIN_SORTED_OUT")) {
        const auto ts = Clock::now();
        std::vector<std::pair<const char *, size_t>> lines;
        for (auto &sh : shards)
            for (auto &t : sh.text) {
                size_t a = 0;
                while (a < t.size()) {
                    size_t b = t.find('\n', a);
                    if (b == std::string::npos) b = t.size();
                    lines.emplace_back(t.data() + a, b - a);
                    a = b + 1;
                }
            }
        auto less = [](const std::pair<const char *, size_t> &x, const std::pair<const char *, size_t> &y) {
            const int c = memcmp(x.first, y.first, std::min(x.second, y.second));
            return c != 0 ? c < 0 : x.second < y.second;
        };
        std::sort(lines.begin(), lines.end(), less);
        std::string out;
        for (size_t i = 0; i < lines.size(); i++) {
            if (i && lines[i].second == lines[i - 1].second && memcmp(lines[i].first, lines[i - 1].first, lines[i].second) == 0) continue;
            out.append(lines[i].first, lines[i].second);
            out.push_back('\n');
            sorted_lines++;
        }
        std::ofstream so(sp);
        so.write(out.data(), (std::streamsize)out.size());
        so.close();
        sorted_ms = ms_since(ts);
        std::cout << "Time elapsed (sorted unique overlap file, " << sorted_lines << " lines): " << sorted_ms << " msec" << std::endl;
    }
    const double wall_s = std::chrono::duration<double>(Clock::now() - t_prog).count();
    std::cout << "Time elapsed (program start to here): " << (long)(wall_s * 1e3 + 0.5) << " msec" << std::endl;

    int rcode = 0;
    uint64_t tiles = 0, cells = 0, cand = 0;
    double dev_ms = 0, sched_ms = 0, init_ms = 0, table_ms = 0;
    int batches = 0;
    for (auto &sh : shards) {
        if (!sh.error.empty()) { fprintf(stderr, "shard %d: %s\n", sh.tid, sh.error.c_str()); rcode = 3; }
        tiles += sh.stats.tiles; cells += sh.stats.cells; cand += sh.cand_fwd + sh.cand_rev;
        dev_ms = std::max(dev_ms, sh.stats.device_ms);
        sched_ms = std::max(sched_ms, sh.stats.wall_ms);
        init_ms = std::max(init_ms, (double)sh.init_ms);
        table_ms = std::max(table_ms, sh.table_kernel_ms);
        batches = std::max(batches, sh.batches);
    }
    {
        std::string tl = "DARWIN_B200_TIMELINE {";
        for (size_t i = 0; i < timeline.size(); i++) {
            char buf[160];
            snprintf(buf, sizeof(buf), "%s\"%s\": %.1f", i ? ", " : "", timeline[i].first, timeline[i].second);
            tl += buf;
        }
        tl += "}";
        puts(tl.c_str());
    }
    // machine-readable summary (one line; everything above mirrors the reference's prints)
    printf("DARWIN_B200_SUMMARY {\"reads\": %zu, \"gpus\": %zu, \"candidates\": %llu, \"tiles\": %llu, \"cells\": %llu, "
           "\"align_phase_ms\": %.3f, \"worker_setup_ms\": %.3f, \"gact_sched_ms\": %.1f, \"gact_kernel_ms\": %.1f, "
           "\"gpu_init_ms\": %.0f, \"seed_table_ms\": %.1f, \"teardown_ms\": %ld, \"chain_batches\": %d, \"wall_s\": %.3f, "
           "\"sorted_unique_lines\": %zu}\n",
           num_reads, shards.size(), (unsigned long long)cand, (unsigned long long)tiles, (unsigned long long)cells, align_us / 1e3,
           setup_us / 1e3, sched_ms, dev_ms, init_ms, table_ms, down_ms, batches, wall_s, sorted_lines);
    delete table;
    return rcode;
}
