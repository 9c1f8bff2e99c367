// gact_engine.cu -- C ABI (include/gact_b200.h) of the B200 GACT tile engine.
//
// Replaces the reference's cuda_host.cu: GPU_init (:193-237), Align_Batch_GPU
// (:23-190), GPU_close (:239-258).  Differences by design: sequences are
// uploaded once and stay 2-bit packed in HBM (the reference re-copies every
// tile's bases per batch, cuda_host.cu:85-163); descriptors/results travel
// through pinned double-buffered slots; errors are returned, never exit().
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <chrono>
#include <new>

#include "gact_common.cuh"
#include "gact_kernels_i32.cuh"
#include "gact_kernels_s16.cuh"
#include "gact_kernels_s16h.cuh"
#include "gact_kernels_it.cuh"
#include "dsoft.cuh"
#include "seed_build.cuh"
#include "host_pool.h"
#include <algorithm>

using namespace gact;

namespace {

thread_local std::string g_create_error;

struct SeqSetHost {
    uint32_t *d_packed = nullptr;             // always (non-empty sets)
    uint8_t *d_bytes = nullptr;               // raw bytes, only when the set holds a byte other than ACGT
    uint32_t *d_exc = nullptr;                // exception bitmap, ditto
    long long len = 0;
    int bits = 0;                             // 2: ACGT only; 8: raw bytes kept beside the packed words
    std::vector<long long> starts;
    std::vector<uint32_t> h_exc;              // host copy of the bitmap (empty: no exception)
    std::vector<uint8_t> seq_exc;             // per sequence: holds an exception
    // any exception among bases [off, off + n)?
    bool range_has_exc(long long off, long long n) const
    {
        if (h_exc.empty() || n <= 0) return false;
        const long long w0 = off >> 5, w1 = (off + n - 1) >> 5;
        for (long long w = w0; w <= w1; w++) {
            uint32_t v = h_exc[(size_t)w];
            if (w == w0) v &= 0xffffffffu << (off & 31);
            if (w == w1 && ((off + n) & 31)) v &= 0xffffffffu >> (32 - ((off + n) & 31));
            if (v) return true;
        }
        return false;
    }
};

struct Slot {
    gact_tile_desc *d_descs = nullptr, *h_descs = nullptr;
    gact_tile_result *d_results = nullptr, *h_results = nullptr;
    uint32_t *d_states = nullptr, *h_states = nullptr;
    EffLen *d_eff = nullptr;
    int *d_first = nullptr, *h_first = nullptr;
    int *d_order = nullptr, *h_order = nullptr;      // tiles by descending reference length (pairs similar tiles)
    int *d_counters = nullptr;            // [0] first pass, [1] main pass, [2] / [3] the same for the raw-byte group
    int n_lut = 0, n_first_lut = 0;       // tiles / first tiles whose query window has no exception: they come first in
                                          // h_order / h_first and run on the score-table kernels, the rest on raw bytes
    int n_narrow = 0;                     // trailing part of the n_lut tiles that runs on the narrow mapping (s16h_narrow)
    int n_it = 0;                         // leading part of the n_lut tiles that goes to the inter-task kernel (multiple of 64)
    int *d_escaped = nullptr;             // tiles the inter-task kernel handed back (band left): redone by the wavefront kernel
    int *h_it_info = nullptr;             // pinned: [0] batches claimed, [1] tiles handed back, of this slot's last launch
    cudaStream_t st_it = nullptr;         // the inter-task kernel's own stream: the wavefront kernels of the same batch fill its tail
    cudaEvent_t ev_it0 = nullptr, ev_it1 = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_done = nullptr, ev_h2d = nullptr, ev_fork = nullptr;
    int n = 0, n_first = 0;
    bool busy = false;
    unsigned long long cells = 0;
};

}  // namespace

struct gact_chain_state;
static void destroy_chain_state(struct gact_engine *e);

// One thread's share of check_descs: tiles [a, b) validated and routed.
struct alignas(128) CheckPart {
    std::vector<int> it, rest;
    std::vector<uint8_t> grp;
    int cnt[4], nf[4], bad;
    unsigned long long cells;
};

static HostPool &host_pool() { static HostPool *p = new HostPool; return *p; }   // lives as long as the process
static int g_copy_threads = 4;

struct gact_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;   // copy streams: descriptor upload / result download overlap the kernels
    cudaStream_t cs[GACT_MAX_INFLIGHT] = {};         // compute streams of the async slots: the kernel of batch k+1
                                                     // fills the SMs that the tail of batch k leaves idle
    gact_params params{};
    KParams kp{};
    int max_tiles = 0;
    int pitch_words = 0;
    int num_sms = 0;
    int variant_req = 0;      // 0 auto, 1 int32, 2 s16x2
    int C = 0;                // columns per lane (int32 kernel)
    bool dir_global = false;
    size_t per_warp_bytes = 0;
    int warps_per_cta = 0, ctas = 0;
    size_t smem_main = 0;
    uint8_t *d_gscratch = nullptr;
    S16HPlan s16h;            // packed s16x2 kernels: two tiles per warp (tile_size <= 320) or one (<= 1024); ok = usable for these params
    S16HPlan s16h_narrow;     // tile kernel, strips of half the width: tiles whose query window is at most half a tile wide
    S16HPlan s16h_lat;        // chain kernel, one tile per warp: used when candidates < chain slots (latency bound)
    // inter-task tile kernel (gact_kernels_it.cuh): one lane per pair of full, non-first tiles
    bool it_ok = false;       // usable for these parameters
    bool it_qs = false;       // query columns staged in shared memory too
    ITGeom it_geom{};
    int it_ctas = 0, it_min_tiles = 0;     // it_ctas: resident groups of 4 warps
    bool narrow_on_it_stream = false;      // GACT_NARROW_STREAM=1
    int it_cta_warps = 1;                  // warps per CTA of the launch (1, 2 or 4; GACT_IT_CTA_WARPS)
    uint8_t *d_it_scratch = nullptr;      // GACT_MAX_INFLIGHT regions: strip edges, row score tables, code words per resident warp
    size_t it_region_bytes = 0, it_edge_b = 0, it_win_b = 0, it_smem = 0;
    SeqSetHost sets[GACT_MAX_SETS];
    Slot slots[GACT_MAX_INFLIGHT];
    int head = 0, tail = 0, inflight = 0;   // async ring
    bool staged = false;
    bool slots_ready = false;
    double last_kernel_ms = -1.0;
    // GACT_HOST_TRACE=1: host-side milliseconds of the tile path, printed when the engine is destroyed
    bool host_trace = false;
    double ht_check = 0, ht_copy_in = 0, ht_launch = 0, ht_sync = 0, ht_copy_out = 0;
    int last_n_it = 0, last_n_escaped = 0;   // inter-task kernel: tiles it took / handed back in the last finished batch
    struct gact_chain_state *chains = nullptr;       // gact_engine_extend_* state (created on first use)
    std::vector<int> scratch_rest;                   // check_descs work arrays (kept to avoid reallocation per batch)
    std::vector<uint8_t> scratch_grp;
    std::vector<CheckPart> check_parts;
    int host_threads = 4;                            // threads of the host-side passes over a large batch (GACT_HOST_THREADS)
    gact_stats stats{};
    std::string err;
};

namespace {

int fail(gact_engine *e, int code, const std::string &msg)
{
    if (e) e->err = msg; else g_create_error = msg;
    return code;
}

#define CU(e, call)                                                                         \
    do {                                                                                    \
        cudaError_t _r = (call);                                                            \
        if (_r != cudaSuccess)                                                              \
            return fail((e), GACT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_r)); \
    } while (0)

// --------------------------------------------------------------------------
// upload: raw bytes -> 2-bit words (ntcoding.cpp:60-72: case folded, anything else 0) + exception bitmap (one bit
// per base that is not one of "ACGT") in one pass; one thread per 32 bases
__global__ void pack2_kernel(const uint8_t *__restrict__ raw, long long len, uint32_t *__restrict__ packed,
                             uint32_t *__restrict__ exc, long long n_exc_words, int *any_exc)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    int bad = 0;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < n_exc_words; w += stride) {
        uint32_t lo = 0, hi = 0, ex = 0;
        const long long b0 = w * 32;
#pragma unroll
        for (int k = 0; k < 32; k++) {
            const long long idx = b0 + k;
            if (idx < len) {
                const int ch = raw[idx];
                uint32_t code = 0;
                switch (ch | 0x20) {
                    case 'a': code = 0; break;
                    case 'c': code = 1; break;
                    case 'g': code = 2; break;
                    case 't': code = 3; break;
                    default: code = 0; break;
                }
                if (ch != 'A' && ch != 'C' && ch != 'G' && ch != 'T') ex |= 1u << k;
                if (k < 16) lo |= code << (2 * k); else hi |= code << (2 * (k - 16));
            }
        }
        packed[2 * w] = lo;
        packed[2 * w + 1] = hi;
        exc[w] = ex;
        bad |= (ex != 0);
    }
    if (bad) atomicOr(any_exc, 1);
}

void free_set(SeqSetHost &s)
{
    if (s.d_packed) cudaFree(s.d_packed);
    if (s.d_bytes) cudaFree(s.d_bytes);
    if (s.d_exc) cudaFree(s.d_exc);
    s = SeqSetHost();
}

void free_slot(Slot &s)
{
    if (s.d_descs) cudaFree(s.d_descs);
    if (s.d_results) cudaFree(s.d_results);
    if (s.d_states) cudaFree(s.d_states);
    if (s.d_eff) cudaFree(s.d_eff);
    if (s.d_first) cudaFree(s.d_first);
    if (s.d_order) cudaFree(s.d_order);
    if (s.h_order) cudaFreeHost(s.h_order);
    if (s.d_counters) cudaFree(s.d_counters);
    if (s.d_escaped) cudaFree(s.d_escaped);
    if (s.h_it_info) cudaFreeHost(s.h_it_info);
    if (s.st_it) { cudaStreamSynchronize(s.st_it); cudaStreamDestroy(s.st_it); }
    if (s.ev_it0) cudaEventDestroy(s.ev_it0);
    if (s.ev_it1) cudaEventDestroy(s.ev_it1);
    if (s.h_descs) cudaFreeHost(s.h_descs);
    if (s.h_results) cudaFreeHost(s.h_results);
    if (s.h_states) cudaFreeHost(s.h_states);
    if (s.h_first) cudaFreeHost(s.h_first);
    if (s.ev_k0) cudaEventDestroy(s.ev_k0);
    if (s.ev_k1) cudaEventDestroy(s.ev_k1);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
    if (s.ev_fork) cudaEventDestroy(s.ev_fork);
    s = Slot();
}

// --------------------------------------------------------------------------
// kernel dispatch tables
typedef void (*main_fn)(const KParams, const gact_tile_desc *, int, const EffLen *, gact_tile_result *,
                        uint32_t *, int, int *, uint8_t *, size_t);
typedef void (*first_fn)(const KParams, const gact_tile_desc *, const int *, int, EffLen *, int *);

main_fn pick_main_i32(int C, bool g)
{
    switch (C) {
        case 8:  return g ? gact_tile_i32_kernel<8, true>  : gact_tile_i32_kernel<8, false>;
        case 10: return g ? gact_tile_i32_kernel<10, true> : gact_tile_i32_kernel<10, false>;
        case 16: return g ? gact_tile_i32_kernel<16, true> : gact_tile_i32_kernel<16, false>;
        default: return g ? gact_tile_i32_kernel<32, true> : gact_tile_i32_kernel<32, false>;
    }
}
first_fn pick_first_i32(int C)
{
    switch (C) {
        case 8:  return gact_first_i32_kernel<8>;
        case 10: return gact_first_i32_kernel<10>;
        case 16: return gact_first_i32_kernel<16>;
        default: return gact_first_i32_kernel<32>;
    }
}
size_t dir_bytes_i32(int C, int rows, int lanes)
{
    switch (C) {
        case 8:  return DirWin<8>::bytes(rows, lanes);
        case 10: return DirWin<10>::bytes(rows, lanes);
        case 16: return DirWin<16>::bytes(rows, lanes);
        default: return DirWin<32>::bytes(rows, lanes);
    }
}

const size_t SMEM_CTA_MAX = 227 * 1024;    // opt-in dynamic shared memory per CTA on sm_100
const size_t SMEM_SM = 228 * 1024;         // per SM, 1 KB reserved per resident CTA

int plan_launch(gact_engine *e)
{
    const int T = e->params.tile_size;
    const int et = e->params.tile_size - e->params.tile_overlap;
    e->C = (T <= 256) ? 8 : (T <= 320) ? 10 : (T <= 512) ? 16 : 32;
    const int C = e->C;
    const int TS = C * 32;
    e->kp.win_rows = (et + 1 < T) ? et + 1 : T;
    int wl = et / C + 2;
    e->kp.win_lanes = wl > 32 ? 32 : wl;
    const size_t dirb = dir_bytes_i32(C, e->kp.win_rows, e->kp.win_lanes);
    const size_t smem_per_warp = TS + dirb;
    // shared-memory window if at least 4 warps fit per SM, else L2-resident scratch
    int fit = (int)((SMEM_SM - 2048) / smem_per_warp);
    if (fit >= 4) {
        e->dir_global = false;
        e->per_warp_bytes = smem_per_warp;
        int wpc = fit > 8 ? 8 : fit;
        // prefer two CTAs per SM when that keeps more warps resident
        int best_w = wpc, best_c = 1, best_total = wpc;
        for (int w = 1; w <= 8; w++) {
            const size_t cta = (size_t)w * smem_per_warp;
            if (cta > SMEM_CTA_MAX) break;
            int c = (int)(SMEM_SM / (cta + 1024));
            if (c > 8) c = 8;
            if (c * w > 32) c = 32 / w;
            if (c * w > best_total || (c * w == best_total && w > best_w)) { best_total = c * w; best_w = w; best_c = c; }
        }
        e->warps_per_cta = best_w;
        e->ctas = best_c * e->num_sms;
        e->smem_main = (size_t)best_w * smem_per_warp;
    } else {
        e->dir_global = true;
        e->per_warp_bytes = dirb;
        e->warps_per_cta = 4;
        const int ctas_per_sm = 4;
        e->ctas = ctas_per_sm * e->num_sms;
        e->smem_main = (size_t)e->warps_per_cta * TS;
        const size_t total = (size_t)e->ctas * e->warps_per_cta * dirb;
        if (cudaMalloc(&e->d_gscratch, total) != cudaSuccess)
            return fail(e, GACT_ERR_NOMEM, "cudaMalloc(direction scratch) failed");
    }
    main_fn f = pick_main_i32(C, e->dir_global);
    CU(e, cudaFuncSetAttribute((const void *)f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_main));
    first_fn ff = pick_first_i32(C);
    CU(e, cudaFuncSetAttribute((const void *)ff, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * TS));
    // experiment knob (not part of the ABI): GACT_S16_WARPS=<resident warps per SM>
    int wps = 0;
    if (const char *w = getenv("GACT_S16_WARPS")) wps = atoi(w);
    if (s16h_make_plan(e->params, e->num_sms, wps, &e->s16h) != 0 ||
        s16h_make_plan(e->params, e->num_sms, wps, &e->s16h_lat, true) != 0)
        return fail(e, GACT_ERR_CUDA, "s16h kernel attribute setup failed");
    // narrow mapping for tiles with a query window of at most half a tile (GACT_NARROW=0 switches it off)
    e->s16h_narrow = S16HPlan();
    if (e->s16h.ok && !(getenv("GACT_NARROW") && atoi(getenv("GACT_NARROW")) == 0) &&
        s16h_make_plan(e->params, e->num_sms, wps, &e->s16h_narrow, false, true) != 0)
        return fail(e, GACT_ERR_CUDA, "s16h kernel attribute setup failed");
    // inter-task kernel: full, non-first tiles, one lane per pair of tiles (experiment knobs: GACT_IT=0 switches it off,
    // GACT_IT_BAND=<half-width of the tagged band>, GACT_IT_MIN=<fewest eligible tiles of a batch worth a launch>)
    e->it_ok = false;
    if (const char *h = getenv("GACT_HOST_TRACE")) e->host_trace = atoi(h) != 0;
    if (const char *h = getenv("GACT_NARROW_STREAM")) e->narrow_on_it_stream = atoi(h) != 0;
    {
        // host-side passes over a large batch (descriptor check + routing, copy-in, copy-out): half of the cores, 2..8
        const int hw = (int)std::thread::hardware_concurrency();
        e->host_threads = std::max(2, std::min(8, hw / 2));
        if (const char *h = getenv("GACT_HOST_THREADS")) e->host_threads = std::max(1, std::min(16, atoi(h)));
        g_copy_threads = e->host_threads;
    }
    const char *itv = getenv("GACT_IT");
    if (e->s16h.ok && e->s16h.lut_ok && T % IT_CS == 0 && T / IT_CS <= IT_MAX_STRIPS && T >= 64 && !(itv && atoi(itv) == 0)) {
        int W = std::max(32, (et + 7) / 8);         // a 15 % error channel drifts ~0.045 et off the diagonal, +- 0.37 sqrt(et)
        if (const char *b = getenv("GACT_IT_BAND")) W = std::max(4, atoi(b));
        e->it_geom = it_geometry(T, et, W);
        int it_per_sm = 4;                                             // 4 CTAs x 4 warps per SM (experiment knob GACT_IT_CTAS)
        if (const char *c = getenv("GACT_IT_CTAS")) it_per_sm = std::max(1, std::min(4, atoi(c)));
        e->it_ctas = it_per_sm * e->num_sms;
        e->it_min_tiles = 64 * 4 * e->num_sms;                         // one warp per SM sub-partition at least
        if (const char *m = getenv("GACT_IT_MIN")) e->it_min_tiles = std::max(64, atoi(m));
        e->it_edge_b = it_edge_bytes(T); e->it_win_b = it_win_bytes(e->it_geom);
        e->it_region_bytes = (size_t)e->it_ctas * 4 * (e->it_edge_b + e->it_win_b);
        if (const char *w = getenv("GACT_IT_CTA_WARPS")) { const int v = atoi(w); if (v == 1 || v == 2 || v == 4) e->it_cta_warps = v; }
        // the query columns join the reference rows in shared memory where that still leaves room for every CTA of an SM
        e->it_qs = (4 * it_smem_per_warp(T, true) + 1024) * (size_t)it_per_sm <= (size_t)227 * 1024;
        if (const char *q = getenv("GACT_IT_QS")) e->it_qs = atoi(q) != 0;
        e->it_smem = 4 * it_smem_per_warp(T, e->it_qs);
        const void *fn = e->it_qs ? (const void *)gact_tile_it_kernel<true> : (const void *)gact_tile_it_kernel<false>;
        e->it_ok = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)e->it_smem) == cudaSuccess;      // scratch is allocated with the batch slots
        if (!e->it_ok) cudaGetLastError();
    }
    if (e->s16h.ok) {
        // One shared-memory carve-out for every tile kernel.  Left to itself the driver gives each kernel the carve-out that
        // fits its own CTAs (164 KB for the inter-task kernel, 132 KB for the wavefront kernels), and CTAs of kernels with
        // different carve-outs cannot share an SM: the wavefront kernels of a batch and the inter-task kernel of the next
        // then take turns instead of filling each other's tails.  Measured on 7-batch e2e steps: 34.2 ms without, 30.6 ms
        // with 72 % (164 KB); the device-resident step is unchanged.  GACT_CARVEOUT=<percent> overrides, -1 leaves the default.
        // Only up to tile_size 320, where it was measured: larger tiles need more shared memory per inter-task warp than
        // 164 KB holds at full occupancy.
        int pct = T <= 320 ? 72 : -1;
        if (const char *c = getenv("GACT_CARVEOUT")) pct = atoi(c);
        if (pct >= 0 && pct <= 100) {
            for (int lut = 0; lut < 2; lut++) {
                cudaFuncSetAttribute((const void *)s16h_pick(e->s16h.CS, e->s16h.lanes, lut != 0), cudaFuncAttributePreferredSharedMemoryCarveout, pct);
                cudaFuncSetAttribute((const void *)s16h_pick_first(e->s16h.CS, e->s16h.lanes, lut != 0), cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            }
            if (e->s16h_narrow.ok)
                cudaFuncSetAttribute((const void *)s16h_pick(e->s16h_narrow.CS, e->s16h_narrow.lanes, true), cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            if (e->it_ok)
                cudaFuncSetAttribute(e->it_qs ? (const void *)gact_tile_it_kernel<true> : (const void *)gact_tile_it_kernel<false>,
                                     cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            cudaGetLastError();
        }
    }
    return GACT_OK;
}

bool use_s16(const gact_engine *e)
{
    if (e->variant_req == 1) return false;
    return e->s16h.ok;
}

int launch_batch(gact_engine *e, Slot &s, cudaStream_t st, int scratch_region)
{
    CU(e, cudaStreamWaitEvent(st, s.ev_h2d, 0));        // descriptors of this batch are on the device
    CU(e, cudaMemsetAsync(s.d_counters, 0, 8 * sizeof(int), st));
    CU(e, cudaEventRecord(s.ev_k0, st));
    const int TS = e->C * 32;
    if (use_s16(e)) {
        // two groups (check_descs): tiles whose query window is free of exceptions score with the one-PRMT table from the
        // packed words; the others compare raw bytes (HSET2 path).  Same kernels otherwise, results land at the tile's index.
        const int n_byte = s.n - s.n_lut, nf_byte = s.n_first - s.n_first_lut;
        if (s.n_it > 0) {
            // full, non-first tiles: inter-task kernel on its own stream, launched first; its CTAs hold all registers of
            // the SMs, so the wavefront kernels below start as its last wave drains and fill that tail
            KParams kp = e->kp;
            kp.s16_bias = e->s16h.bias;
            kp.one = 1;
            const int n_batches = s.n_it / 64;
            // one-warp CTAs: a warp that finds no further group of 64 tiles retires at once and hands its registers to
            // the kernels queued behind (the next batch's, the wavefront kernels), instead of idling until the slowest
            // warp of a 4-warp CTA is done
            const int cw = e->it_cta_warps;
            int grid = (n_batches + cw - 1) / cw;
            if (grid > e->it_ctas * 4 / cw) grid = e->it_ctas * 4 / cw;
            uint8_t *base = e->d_it_scratch + (size_t)scratch_region * e->it_region_bytes;
            const size_t warps = (size_t)e->it_ctas * 4;
            uint2 *edge = reinterpret_cast<uint2 *>(base);
            uint32_t *win = reinterpret_cast<uint32_t *>(base + warps * e->it_edge_b);
            CU(e, cudaEventRecord(s.ev_it0, st));
            CU(e, cudaStreamWaitEvent(s.st_it, s.ev_it0, 0));
            auto kern = e->it_qs ? gact_tile_it_kernel<true> : gact_tile_it_kernel<false>;
            kern<<<grid, 32 * cw, e->it_smem / 4 * cw, s.st_it>>>(kp, e->it_geom, s.d_descs, s.d_order, n_batches, s.d_results, s.d_states,
                                                     e->pitch_words, s.d_counters + 4, s.d_escaped, edge, win, e->it_win_b / 4);
            e->stats.kernel_launches++;
            if (s.n_narrow > 0 && e->narrow_on_it_stream) {
                // the narrow group needs nothing from the other kernels of the batch: behind the inter-task kernel on its
                // stream it runs while the main stream's kernels do, instead of adding a launch tail of its own there
                s16h_launch(e->s16h_narrow, e->kp, s.d_descs, s.d_order + s.n_lut - s.n_narrow, s.n_narrow, s.d_eff, s.d_results,
                            s.d_states, e->pitch_words, s.d_counters + 7, s.st_it, scratch_region, true);
                e->stats.kernel_launches++;
            }
            CU(e, cudaEventRecord(s.ev_it1, s.st_it));
        }
        if (s.n_first_lut > 0) {
            s16h_launch_first(e->s16h, e->kp, s.d_descs, s.d_first, s.n_first_lut, s.d_eff, s.d_counters + 0, st, true);
            e->stats.kernel_launches++;
        }
        if (nf_byte > 0) {
            s16h_launch_first(e->s16h, e->kp, s.d_descs, s.d_first + s.n_first_lut, nf_byte, s.d_eff, s.d_counters + 2, st, false);
            e->stats.kernel_launches++;
        }
        if (s.n_lut - s.n_it - s.n_narrow > 0) {
            s16h_launch(e->s16h, e->kp, s.d_descs, s.d_order + s.n_it, s.n_lut - s.n_it - s.n_narrow, s.d_eff, s.d_results,
                        s.d_states, e->pitch_words, s.d_counters + 1, st, scratch_region, true);
            e->stats.kernel_launches++;
        }
        if (s.n_narrow > 0 && !(e->narrow_on_it_stream && s.n_it > 0)) {
            s16h_launch(e->s16h_narrow, e->kp, s.d_descs, s.d_order + s.n_lut - s.n_narrow, s.n_narrow, s.d_eff, s.d_results,
                        s.d_states, e->pitch_words, s.d_counters + 7, st, scratch_region, true);
            e->stats.kernel_launches++;
        }
        if (n_byte > 0) {
            s16h_launch(e->s16h, e->kp, s.d_descs, s.d_order + s.n_lut, n_byte, s.d_eff, s.d_results, s.d_states, e->pitch_words,
                        s.d_counters + 3, st, scratch_region, false);
            e->stats.kernel_launches++;
        }
        if (s.n_it > 0) {
            // tiles whose traceback left the band: redone with the full window; their number stays on the device
            CU(e, cudaStreamWaitEvent(st, s.ev_it1, 0));
            const int bound = std::min(s.n_it, e->s16h.ctas * e->s16h.warps_per_cta * e->s16h.tpw());
            s16h_launch(e->s16h, e->kp, s.d_descs, s.d_escaped, bound, s.d_eff, s.d_results, s.d_states, e->pitch_words,
                        s.d_counters + 6, st, scratch_region, true, s.d_counters + 5);
            e->stats.kernel_launches++;
        }
    } else {
        if (s.n_first > 0) {
            first_fn ff = pick_first_i32(e->C);
            int ctas = e->num_sms * 4;
            int need = (s.n_first + 7) / 8;
            if (need < ctas) ctas = need;
            ff<<<ctas, 256, 8 * TS, st>>>(e->kp, s.d_descs, s.d_first, s.n_first, s.d_eff, s.d_counters + 0);
            e->stats.kernel_launches++;
        }
        main_fn f = pick_main_i32(e->C, e->dir_global);
        int ctas = e->ctas;
        int need = (s.n + e->warps_per_cta - 1) / e->warps_per_cta;
        if (need < ctas) ctas = need;
        f<<<ctas, e->warps_per_cta * 32, e->smem_main, st>>>(e->kp, s.d_descs, s.n, s.d_eff, s.d_results,
                                                             s.d_states, e->pitch_words, s.d_counters + 1,
                                                             e->d_gscratch, e->per_warp_bytes);
        e->stats.kernel_launches++;
    }
    CU(e, cudaGetLastError());
    CU(e, cudaEventRecord(s.ev_k1, st));
    return GACT_OK;
}

int check_descs(gact_engine *e, int n, const gact_tile_desc *descs, Slot &s)
{
    const int T = e->params.tile_size;
    const bool table_ok = use_s16(e) && e->s16h.lut_ok;
    // group 0: inter-task kernel (full, non-first tiles of the score-table group), 1: score-table wavefront kernels,
    // 2: raw-byte wavefront kernels, 3: score-table wavefront kernel in its narrow mapping (non-first tiles whose query
    // window fits strips of half the width: the same rows at half the work per wavefront step).  One pass over the descriptors, on a few threads for a large batch (10 ns per tile
    // on one core is 10 ms per Mi tiles, as long as the kernels of a 256 Ki batch take): inter-task tiles go straight
    // into h_order (their order does not matter), the others are collected and counting-sorted afterwards (they are
    // the minority of a large batch).
    const bool it_on = table_ok && e->it_ok && e->d_it_scratch && s.d_escaped;
    const int narrow_cols = (table_ok && e->s16h_narrow.ok) ? e->s16h_narrow.cols() : 0;
    const int parts = n >= (1 << 16) ? std::max(1, std::min(8, e->host_threads)) : 1;
    if ((int)e->check_parts.size() < parts) e->check_parts.resize((size_t)parts);
    auto work = [&](int p) {
        // everything a tile touches is local to the thread (the parts sit next to each other in memory)
        CheckPart c;
        {
            CheckPart &mine = e->check_parts[(size_t)p];
            c.it.swap(mine.it); c.rest.swap(mine.rest); c.grp.swap(mine.grp);     // keep the capacity of the last batch
        }
        c.it.clear(); c.rest.clear(); c.grp.clear();
        for (int g = 0; g < 4; g++) c.cnt[g] = c.nf[g] = 0;
        c.cells = 0; c.bad = -1;
        const int a = (int)((long long)n * p / parts), b = (int)((long long)n * (p + 1) / parts);
        for (int t = a; t < b && c.bad < 0; t++) {
            const gact_tile_desc &d = descs[t];
            if (d.ref_set >= GACT_MAX_SETS || d.query_set >= GACT_MAX_SETS || d.ref_len < 0 || d.query_len < 0 ||
                d.ref_len > T || d.query_len > T || d.ref_off < 0 || d.query_off < 0 ||
                d.ref_off + d.ref_len > e->sets[d.ref_set].len || d.query_off + d.query_len > e->sets[d.query_set].len) {
                c.bad = t;
                break;
            }
            int g = (table_ok && !e->sets[d.query_set].range_has_exc(d.query_off, d.query_len)) ? 1 : 2;
            if (g == 1 && it_on && !d.first && d.ref_len == T && d.query_len == T && e->sets[d.ref_set].h_exc.empty()) g = 0;
            else if (g == 1 && narrow_cols > 0 && !d.first && d.query_len <= narrow_cols) g = 3;
            c.cnt[g]++;
            if (d.first) c.nf[g]++;
            c.cells += (unsigned long long)d.ref_len * (unsigned long long)d.query_len;
            if (g == 0) c.it.push_back(t);
            else { c.rest.push_back(t); c.grp.push_back((uint8_t)g); }
        }
        CheckPart &mine = e->check_parts[(size_t)p];
        mine.it.swap(c.it); mine.rest.swap(c.rest); mine.grp.swap(c.grp);
        for (int g = 0; g < 4; g++) { mine.cnt[g] = c.cnt[g]; mine.nf[g] = c.nf[g]; }
        mine.cells = c.cells; mine.bad = c.bad;
    };
    host_pool().run(parts, work);
    std::vector<int> &rest = e->scratch_rest;
    std::vector<uint8_t> &rgrp = e->scratch_grp;
    rest.clear(); rgrp.clear();
    unsigned long long cells = 0;
    int cnt[4] = {0, 0, 0, 0}, nf[4] = {0, 0, 0, 0};
    int pos_it = 0;
    for (int p = 0; p < parts; p++) {
        const CheckPart &c = e->check_parts[(size_t)p];
        if (c.bad >= 0) return fail(e, GACT_ERR_ARG, "tile descriptor " + std::to_string(c.bad) + " out of range");
        for (int g = 0; g < 4; g++) { cnt[g] += c.cnt[g]; nf[g] += c.nf[g]; }
        cells += c.cells;
        if (!c.it.empty()) memcpy(s.h_order + pos_it, c.it.data(), c.it.size() * sizeof(int));
        pos_it += (int)c.it.size();
        rest.insert(rest.end(), c.rest.begin(), c.rest.end());
        rgrp.insert(rgrp.end(), c.grp.begin(), c.grp.end());
    }
    // the inter-task kernel takes whole warps of 64 tiles and only batches that fill the GPU; the rest joins group 1
    int n_it = (cnt[0] / 64) * 64;
    if (n_it < e->it_min_tiles) n_it = 0;
    for (int k = n_it; k < cnt[0]; k++) { rest.push_back(s.h_order[k]); rgrp.push_back(1); }
    cnt[1] += cnt[0] - n_it;
    cnt[0] = n_it;
    if (cnt[3] > 0 && cnt[3] < 64) {                 // not worth a launch of its own
        for (auto &g : rgrp) if (g == 3) g = 1;
        cnt[1] += cnt[3];
        cnt[3] = 0;
    }
    s.n_it = n_it;
    s.n_lut = cnt[0] + cnt[1] + cnt[3];
    s.n_narrow = cnt[3];
    s.n_first_lut = nf[1];
    s.n_first = nf[1] + nf[2];
    s.cells = cells;
    int fpos[4] = {0, 0, nf[1], 0};
    for (size_t k = 0; k < rest.size(); k++) if (descs[rest[k]].first) s.h_first[fpos[rgrp[k]]++] = rest[k];
    // h_order: [inter-task tiles][score-table tiles][narrow score-table tiles][raw-byte tiles]; the wavefront groups are
    // counting-sorted by reference length, longest first: the two tiles a warp aligns side by side then have the same
    // number of wavefront steps, and the long tiles start first
    static const int slot_of_group[4] = {0, 0, 2, 1};
    std::vector<int> start(3 * ((size_t)T + 2) + 1, 0);
    auto bucket = [&](size_t k) { return (size_t)slot_of_group[rgrp[k]] * ((size_t)T + 1) + (size_t)(T - descs[rest[k]].ref_len); };
    for (size_t k = 0; k < rest.size(); k++) start[bucket(k) + 1]++;
    for (size_t k = 1; k < start.size(); k++) start[k] += start[k - 1];
    for (size_t k = 0; k < rest.size(); k++) s.h_order[n_it + start[bucket(k)]++] = rest[k];
    return GACT_OK;
}

void par_memcpy(void *dst, const void *src, size_t bytes);
static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int enqueue(gact_engine *e, Slot &s, int n, const gact_tile_desc *descs)
{
    if (n < 0 || n > e->max_tiles || (n > 0 && !descs)) return fail(e, GACT_ERR_ARG, "bad tile count");
    const double t0 = e->host_trace ? now_ms() : 0;
    int rc = check_descs(e, n, descs, s);
    if (rc) return rc;
    s.n = n;
    if (n == 0) return GACT_OK;
    const double t1 = e->host_trace ? now_ms() : 0;
    par_memcpy(s.h_descs, descs, (size_t)n * sizeof(gact_tile_desc));
    if (e->host_trace) { e->ht_check += t1 - t0; e->ht_copy_in += now_ms() - t1; }
    // upload on the H2D copy stream; the compute stream waits for it, so the copy of batch k+1
    // overlaps the kernels of batch k
    CU(e, cudaMemcpyAsync(s.d_descs, s.h_descs, (size_t)n * sizeof(gact_tile_desc), cudaMemcpyHostToDevice, e->s_h2d));
    if (s.n_first)
        CU(e, cudaMemcpyAsync(s.d_first, s.h_first, (size_t)s.n_first * sizeof(int), cudaMemcpyHostToDevice, e->s_h2d));
    CU(e, cudaMemcpyAsync(s.d_order, s.h_order, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, e->s_h2d));
    CU(e, cudaEventRecord(s.ev_h2d, e->s_h2d));
    e->stats.h2d_bytes += (double)n * (sizeof(gact_tile_desc) + sizeof(int)) + (double)s.n_first * sizeof(int);
    return GACT_OK;
}

int download(gact_engine *e, Slot &s, bool want_states)
{
    if (s.n == 0) return GACT_OK;
    // download on the D2H copy stream once the kernels of this batch are done (ev_k1), so the
    // next batch's kernels do not wait for the copy
    CU(e, cudaStreamWaitEvent(e->s_d2h, s.ev_k1, 0));
    CU(e, cudaMemcpyAsync(s.h_results, s.d_results, (size_t)s.n * sizeof(gact_tile_result), cudaMemcpyDeviceToHost, e->s_d2h));
    if (s.n_it > 0) CU(e, cudaMemcpyAsync(s.h_it_info, s.d_counters + 4, 2 * sizeof(int), cudaMemcpyDeviceToHost, e->s_d2h));
    e->stats.d2h_bytes += (double)s.n * sizeof(gact_tile_result);
    if (want_states) {
        CU(e, cudaMemcpyAsync(s.h_states, s.d_states, (size_t)s.n * e->pitch_words * 4, cudaMemcpyDeviceToHost, e->s_d2h));
        e->stats.d2h_bytes += (double)s.n * e->pitch_words * 4;
    }
    return GACT_OK;
}

// copy-out of a large batch on a few threads (one core moves ~8 GB/s; 1 Mi tiles are 134 MB of results + states)
void par_memcpy(void *dst, const void *src, size_t bytes)
{
    const size_t MIN = (size_t)2 << 20;
    if (bytes < 2 * MIN) { memcpy(dst, src, bytes); return; }
    const int parts = (int)std::min<size_t>((size_t)g_copy_threads, bytes / MIN);
    const size_t per = ((bytes / parts) + 63) & ~(size_t)63;
    host_pool().run(parts, [=](int p) {
        const size_t a = per * p, b = (p == parts - 1) ? bytes : std::min(bytes, per * (p + 1));
        if (a < b) memcpy((char *)dst + a, (const char *)src + a, b - a);
    });
}

int finish(gact_engine *e, Slot &s, gact_tile_result *results, uint32_t *packed_states, bool states_copied)
{
    const double t0 = e->host_trace ? now_ms() : 0;
    CU(e, cudaEventSynchronize(s.ev_done));
    const double t1 = e->host_trace ? now_ms() : 0;
    if (e->host_trace) e->ht_sync += t1 - t0;
    if (s.n > 0) {
        float ms = 0.f;
        CU(e, cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1));
        e->last_kernel_ms = ms;
        e->stats.kernel_ms += ms;
        e->last_n_it = s.n_it;
        e->last_n_escaped = s.n_it > 0 ? s.h_it_info[1] : 0;
        if (results) par_memcpy(results, s.h_results, (size_t)s.n * sizeof(gact_tile_result));
        if (packed_states && states_copied) par_memcpy(packed_states, s.h_states, (size_t)s.n * e->pitch_words * 4);
        if (e->host_trace) e->ht_copy_out += now_ms() - t1;
    }
    e->stats.tiles += s.n;
    e->stats.cells += s.cells;
    e->stats.first_tiles += s.n_first;
    e->stats.batches++;
    return GACT_OK;
}

// The batch slots of the tile path (device + pinned host copies of descriptors, results, states) are allocated at the first
// tile batch: a caller that only extends whole candidates on the device (gact_engine_extend) never pays for them.
int ensure_slots(gact_engine *e)
{
    if (e->slots_ready) return GACT_OK;
    if (e->it_ok && e->max_tiles >= e->it_min_tiles && !e->d_it_scratch) {
        if (cudaMalloc(&e->d_it_scratch, (size_t)GACT_MAX_INFLIGHT * e->it_region_bytes) != cudaSuccess) {
            cudaGetLastError();
            e->d_it_scratch = nullptr;
            e->it_ok = false;                 // not enough memory for the scratch: the wavefront kernels take every tile
        }
    } else if (e->it_ok && e->max_tiles < e->it_min_tiles) {
        e->it_ok = false;                     // batches of this engine are too small to fill the GPU with one lane per tile pair
    }
    for (int k = 0; k < GACT_MAX_INFLIGHT; k++) {
        Slot &s = e->slots[k];
        const size_t n = (size_t)e->max_tiles;
        CU(e, cudaMalloc(&s.d_descs, n * sizeof(gact_tile_desc)));
        CU(e, cudaMalloc(&s.d_results, n * sizeof(gact_tile_result)));
        CU(e, cudaMalloc(&s.d_states, n * e->pitch_words * 4));
        CU(e, cudaMalloc(&s.d_eff, n * sizeof(EffLen)));
        CU(e, cudaMalloc(&s.d_first, n * sizeof(int)));
        CU(e, cudaMalloc(&s.d_order, n * sizeof(int)));
        CU(e, cudaMallocHost(&s.h_order, n * sizeof(int)));
        CU(e, cudaMalloc(&s.d_counters, 8 * sizeof(int)));
        if (e->it_ok) CU(e, cudaMalloc(&s.d_escaped, n * sizeof(int)));
        CU(e, cudaMallocHost(&s.h_it_info, 2 * sizeof(int)));
        if (e->it_ok) {
            CU(e, cudaStreamCreateWithFlags(&s.st_it, cudaStreamNonBlocking));
            CU(e, cudaEventCreateWithFlags(&s.ev_it0, cudaEventDisableTiming));
            CU(e, cudaEventCreateWithFlags(&s.ev_it1, cudaEventDisableTiming));
        }
        s.h_it_info[0] = s.h_it_info[1] = 0;
        CU(e, cudaMallocHost(&s.h_descs, n * sizeof(gact_tile_desc)));
        CU(e, cudaMallocHost(&s.h_results, n * sizeof(gact_tile_result)));
        CU(e, cudaMallocHost(&s.h_states, n * e->pitch_words * 4));
        CU(e, cudaMallocHost(&s.h_first, n * sizeof(int)));
        CU(e, cudaEventCreate(&s.ev_k0));
        CU(e, cudaEventCreate(&s.ev_k1));
        CU(e, cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
        CU(e, cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
        CU(e, cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming));
    }
    e->slots_ready = true;
    return GACT_OK;
}

}  // namespace

// ===========================================================================
extern "C" {

int gact_abi_version(void) { return GACT_B200_ABI_VERSION; }

const char *gact_status_string(int status)
{
    switch (status) {
        case GACT_OK: return "ok";
        case GACT_ERR_ARG: return "bad argument";
        case GACT_ERR_CUDA: return "CUDA error";
        case GACT_ERR_NOMEM: return "out of memory";
        case GACT_ERR_STATE: return "call sequence error";
        case GACT_ERR_NODEVICE: return "no CUDA device";
        default: return "unknown status";
    }
}

int gact_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return -1; }
    return n;
}

const char *gact_last_error(const gact_engine *e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int gact_engine_create(gact_engine **out, int device, const gact_params *p, int max_tiles, void *stream)
{
    if (!out || !p) return fail(nullptr, GACT_ERR_ARG, "null argument");
    *out = nullptr;
    if (p->tile_size < 1 || p->tile_size > GACT_MAX_TILE_SIZE || p->tile_overlap < 0 ||
        p->tile_overlap >= p->tile_size)
        return fail(nullptr, GACT_ERR_ARG, "tile_size/tile_overlap out of range");
    if (p->gap_open > 0 || p->gap_extend > 0)
        return fail(nullptr, GACT_ERR_ARG, "gap_open and gap_extend must be <= 0");
    if (abs(p->match) > 1024 || abs(p->mismatch) > 1024 || p->gap_open < -1024 || p->gap_extend < -1024)
        return fail(nullptr, GACT_ERR_ARG, "scores out of range (|score| <= 1024)");
    if (max_tiles < 1) return fail(nullptr, GACT_ERR_ARG, "max_tiles_per_batch must be >= 1");
    int ndev = gact_device_count();
    if (ndev <= 0) return fail(nullptr, GACT_ERR_NODEVICE, "no CUDA device available");
    if (device < 0 || device >= ndev) return fail(nullptr, GACT_ERR_ARG, "device index out of range");

    gact_engine *e = new (std::nothrow) gact_engine();
    if (!e) return fail(nullptr, GACT_ERR_NOMEM, "host allocation failed");
    e->device = device;
    e->params = *p;
    e->max_tiles = max_tiles;
    int rc = GACT_OK;
#define CK(call)                                                                                 \
    do {                                                                                         \
        cudaError_t _r = (call);                                                                 \
        if (_r != cudaSuccess) {                                                                 \
            rc = fail(nullptr, GACT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_r)); \
            goto bad;                                                                            \
        }                                                                                        \
    } while (0)
    {
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) {
            rc = fail(nullptr, GACT_ERR_NODEVICE, "device is not sm_100 class (this library is built for sm_100a only)");
            goto bad;
        }
        e->num_sms = prop.multiProcessorCount;
        if (stream) { e->stream = (cudaStream_t)stream; e->owns_stream = false; }
        else { CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)); e->owns_stream = true; }
        CK(cudaStreamCreateWithFlags(&e->s_h2d, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&e->s_d2h, cudaStreamNonBlocking));
        for (int k = 0; k < GACT_MAX_INFLIGHT; k++) CK(cudaStreamCreateWithFlags(&e->cs[k], cudaStreamNonBlocking));

        const int et = p->tile_size - p->tile_overlap;
        e->pitch_words = (2 * et + 15) / 16 + 1;
        e->kp.match = p->match; e->kp.mismatch = p->mismatch;
        e->kp.gap_open = p->gap_open; e->kp.gap_extend = p->gap_extend;
        e->kp.et = et; e->kp.tile_size = p->tile_size;
        for (int i = 0; i < GACT_MAX_SETS; i++) e->kp.sets[i] = SeqSetDev{nullptr, nullptr, nullptr, 0};

        rc = plan_launch(e);
        if (rc) { g_create_error = e->err; goto bad; }

    }
#undef CK
    *out = e;
    return GACT_OK;
bad:
    gact_engine_destroy(e);
    return rc;
}

void gact_engine_destroy(gact_engine *e)
{
    if (!e) return;
    if (e->host_trace)
        fprintf(stderr, "GACT_HOST_TRACE tile path, host ms: check_descs %.1f  copy-in %.1f  launch+enqueue %.1f  wait(sync) %.1f  copy-out %.1f  "
                "(%llu batches, %llu tiles)\n", e->ht_check, e->ht_copy_in, e->ht_launch, e->ht_sync, e->ht_copy_out,
                (unsigned long long)e->stats.batches, (unsigned long long)e->stats.tiles);
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (int k = 0; k < GACT_MAX_INFLIGHT; k++) if (e->cs[k]) cudaStreamSynchronize(e->cs[k]);
    if (e->s_d2h) cudaStreamSynchronize(e->s_d2h);
    for (int i = 0; i < GACT_MAX_SETS; i++) free_set(e->sets[i]);
    for (int k = 0; k < GACT_MAX_INFLIGHT; k++) free_slot(e->slots[k]);
    if (e->d_gscratch) cudaFree(e->d_gscratch);
    if (e->d_it_scratch) cudaFree(e->d_it_scratch);
    s16h_free_plan(&e->s16h);
    s16h_free_plan(&e->s16h_lat);
    s16h_free_plan(&e->s16h_narrow);
    destroy_chain_state(e);
    if (e->s_h2d) { cudaStreamSynchronize(e->s_h2d); cudaStreamDestroy(e->s_h2d); }
    if (e->s_d2h) { cudaStreamSynchronize(e->s_d2h); cudaStreamDestroy(e->s_d2h); }
    for (int k = 0; k < GACT_MAX_INFLIGHT; k++) if (e->cs[k]) { cudaStreamSynchronize(e->cs[k]); cudaStreamDestroy(e->cs[k]); }
    if (e->owns_stream && e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int gact_engine_upload(gact_engine *e, int set, int64_t n_seqs, const char *const *seqs, const int64_t *lens)
{
    if (!e) return GACT_ERR_ARG;
    if (set < 0 || set >= GACT_MAX_SETS || n_seqs < 0 || (n_seqs > 0 && (!seqs || !lens)))
        return fail(e, GACT_ERR_ARG, "bad upload arguments");
    if (e->inflight || e->staged) return fail(e, GACT_ERR_STATE, "upload while batches are outstanding");
    CU(e, cudaSetDevice(e->device));
    CU(e, cudaStreamSynchronize(e->stream));
    SeqSetHost &s = e->sets[set];
    free_set(s);
    e->kp.sets[set] = SeqSetDev{nullptr, nullptr, nullptr, 0};
    // the set becomes visible (len, starts, device pointers) only once every step below has succeeded: a failed
    // upload leaves it empty, so descriptor validation rejects tiles that would address it
    long long total = 0;
    std::vector<long long> starts((size_t)n_seqs + 1);
    for (int64_t i = 0; i < n_seqs; i++) {
        if (lens[i] < 0 || (lens[i] > 0 && !seqs[i])) return fail(e, GACT_ERR_ARG, "bad sequence in upload");
        starts[(size_t)i] = total;
        total += lens[i];
    }
    starts[(size_t)n_seqs] = total;
    if (total == 0) { s.starts = starts; s.len = 0; s.bits = 0; s.seq_exc.assign((size_t)n_seqs, 0); return GACT_OK; }

    // stage through pinned memory in chunks, concatenating on the device
    uint8_t *d_raw = nullptr;
    const long long padded = (total + 63) & ~63LL;
    if (cudaMalloc(&d_raw, (size_t)padded) != cudaSuccess) { cudaGetLastError(); return fail(e, GACT_ERR_NOMEM, "cudaMalloc(raw bases) failed"); }
    const size_t CH = 32u << 20;
    uint8_t *h_stage = nullptr;
    if (cudaMallocHost(&h_stage, CH) != cudaSuccess) { cudaGetLastError(); cudaFree(d_raw); return fail(e, GACT_ERR_NOMEM, "cudaMallocHost(stage) failed"); }
    long long done = 0;        // bases already sent
    size_t fill = 0;
    int rc = GACT_OK;
    auto flush = [&]() -> int {
        if (!fill) return GACT_OK;
        cudaError_t r = cudaMemcpyAsync(d_raw + done, h_stage, fill, cudaMemcpyHostToDevice, e->stream);
        if (r == cudaSuccess) r = cudaStreamSynchronize(e->stream);
        if (r != cudaSuccess) return fail(e, GACT_ERR_CUDA, std::string("upload copy: ") + cudaGetErrorString(r));
        done += (long long)fill;
        fill = 0;
        return GACT_OK;
    };
    for (int64_t i = 0; i < n_seqs && rc == GACT_OK; i++) {
        long long off = 0;
        while (off < lens[i]) {
            size_t take = (size_t)std::min<long long>(lens[i] - off, (long long)(CH - fill));
            memcpy(h_stage + fill, seqs[i] + off, take);
            fill += take; off += (long long)take;
            if (fill == CH && (rc = flush()) != GACT_OK) break;
        }
    }
    if (rc == GACT_OK) rc = flush();
    cudaFreeHost(h_stage);
    if (rc != GACT_OK) { cudaFree(d_raw); return rc; }
    e->stats.h2d_bytes += (double)total;

    const long long n_exc_words = (total + 31) / 32;
    const long long n_words = 2 * n_exc_words + 2;       // pad words: tiles may peek past the end
    uint32_t *d_packed = nullptr, *d_exc = nullptr;
    int *d_flag = nullptr;
    if (cudaMalloc(&d_packed, (size_t)n_words * 4) != cudaSuccess || cudaMalloc(&d_exc, (size_t)(n_exc_words + 2) * 4) != cudaSuccess ||
        cudaMalloc(&d_flag, sizeof(int)) != cudaSuccess) {
        cudaGetLastError(); cudaFree(d_raw); if (d_packed) cudaFree(d_packed); if (d_exc) cudaFree(d_exc);
        return fail(e, GACT_ERR_NOMEM, "cudaMalloc(packed bases) failed");
    }
    int h_flag = 0;
    cudaMemsetAsync(d_flag, 0, sizeof(int), e->stream);
    cudaMemsetAsync(d_packed, 0, (size_t)n_words * 4, e->stream);
    cudaMemsetAsync(d_exc, 0, (size_t)(n_exc_words + 2) * 4, e->stream);
    int blocks = (int)std::min<long long>((n_exc_words + 255) / 256, (long long)e->num_sms * 8);
    pack2_kernel<<<blocks, 256, 0, e->stream>>>(d_raw, total, d_packed, d_exc, n_exc_words, d_flag);
    cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, e->stream);
    cudaError_t r = cudaStreamSynchronize(e->stream);
    cudaFree(d_flag);
    if (r != cudaSuccess) { cudaFree(d_raw); cudaFree(d_packed); cudaFree(d_exc); return fail(e, GACT_ERR_CUDA, std::string("pack kernel: ") + cudaGetErrorString(r)); }
    std::vector<uint32_t> h_exc;
    std::vector<uint8_t> seq_exc((size_t)n_seqs, 0);
    if (h_flag) {
        // sets with exceptions keep the bitmap (device + host copy) and the raw bytes for the byte-comparing kernels
        h_exc.resize((size_t)n_exc_words + 2, 0);
        r = cudaMemcpy(h_exc.data(), d_exc, (size_t)n_exc_words * 4, cudaMemcpyDeviceToHost);
        if (r != cudaSuccess) { cudaFree(d_raw); cudaFree(d_packed); cudaFree(d_exc); return fail(e, GACT_ERR_CUDA, std::string("exception bitmap copy: ") + cudaGetErrorString(r)); }
        s.d_bytes = d_raw; s.d_exc = d_exc; s.bits = 8;
        e->stats.d2h_bytes += (double)n_exc_words * 4;
    } else {
        cudaFree(d_raw); cudaFree(d_exc); s.bits = 2;
    }
    s.d_packed = d_packed;
    s.h_exc.swap(h_exc);
    s.starts.swap(starts);
    s.len = total;
    if (h_flag)
        for (int64_t i = 0; i < n_seqs; i++) seq_exc[(size_t)i] = s.range_has_exc(s.starts[(size_t)i], lens[i]) ? 1 : 0;
    s.seq_exc.swap(seq_exc);
    e->kp.sets[set] = SeqSetDev{s.d_packed, s.d_bytes, s.d_exc, s.len};
    return GACT_OK;
}

int64_t gact_engine_seq_start(const gact_engine *e, int set, int64_t i)
{
    if (!e || set < 0 || set >= GACT_MAX_SETS) return -1;
    const SeqSetHost &s = e->sets[set];
    if (i < 0 || (size_t)i >= s.starts.size()) return -1;
    return s.starts[(size_t)i];
}
int64_t gact_engine_set_length(const gact_engine *e, int set)
{
    if (!e || set < 0 || set >= GACT_MAX_SETS) return -1;
    return e->sets[set].len;
}
int gact_engine_seq_has_exceptions(const gact_engine *e, int set, int64_t i)
{
    if (!e || set < 0 || set >= GACT_MAX_SETS) return -1;
    const SeqSetHost &s = e->sets[set];
    if (i < 0 || (size_t)i + 1 >= s.starts.size()) return -1;
    return (!s.seq_exc.empty() && s.seq_exc[(size_t)i]) ? 1 : 0;
}
int gact_engine_set_bits(const gact_engine *e, int set)
{
    if (!e || set < 0 || set >= GACT_MAX_SETS) return -1;
    return e->sets[set].bits;
}
int gact_engine_states_pitch_words(const gact_engine *e) { return e ? e->pitch_words : -1; }
int gact_engine_max_tiles(const gact_engine *e) { return e ? e->max_tiles : -1; }

int gact_engine_submit(gact_engine *e, int n, const gact_tile_desc *descs)
{
    if (!e) return GACT_ERR_ARG;
    { CU(e, cudaSetDevice(e->device)); int rs = ensure_slots(e); if (rs) return rs; }
    if (e->staged) return fail(e, GACT_ERR_STATE, "submit while a staged batch is pending");
    if (e->inflight >= GACT_MAX_INFLIGHT) return fail(e, GACT_ERR_STATE, "GACT_MAX_INFLIGHT batches already in flight");
    CU(e, cudaSetDevice(e->device));
    Slot &s = e->slots[e->head];
    // The two slots launch on their own streams, ordered after the work already enqueued on the
    // caller's stream.  Kernels that share one scratch area (int32 / one-tile-per-warp variants) stay
    // on one stream.
    const bool overlap = use_s16(e);
    cudaStream_t st = e->cs[overlap ? e->head : 0];
    CU(e, cudaEventRecord(s.ev_fork, e->stream));
    CU(e, cudaStreamWaitEvent(st, s.ev_fork, 0));
    int rc = enqueue(e, s, n, descs);
    if (rc) return rc;
    const double tl = e->host_trace ? now_ms() : 0;
    if (n > 0) {
        rc = launch_batch(e, s, st, overlap ? e->head : 0);
        if (rc) return rc;
        rc = download(e, s, true);
        if (rc) return rc;
    }
    CU(e, cudaEventRecord(s.ev_done, n > 0 ? e->s_d2h : st));
    if (e->host_trace) e->ht_launch += now_ms() - tl;
    s.busy = true;
    e->head = (e->head + 1) % GACT_MAX_INFLIGHT;
    e->inflight++;
    return GACT_OK;
}

int gact_engine_wait(gact_engine *e, gact_tile_result *results, uint32_t *packed_states)
{
    if (!e) return GACT_ERR_ARG;
    if (e->inflight == 0) return fail(e, GACT_ERR_STATE, "wait without submit");
    CU(e, cudaSetDevice(e->device));
    Slot &s = e->slots[e->tail];
    int rc = finish(e, s, results, packed_states, true);
    s.busy = false;
    e->tail = (e->tail + 1) % GACT_MAX_INFLIGHT;
    e->inflight--;
    return rc;
}

int gact_engine_wait_view(gact_engine *e, int *n, const gact_tile_result **results, const uint32_t **packed_states)
{
    if (!e) return GACT_ERR_ARG;
    if (e->inflight == 0) return fail(e, GACT_ERR_STATE, "wait without submit");
    CU(e, cudaSetDevice(e->device));
    Slot &s = e->slots[e->tail];
    int rc = finish(e, s, nullptr, nullptr, false);
    if (n) *n = s.n;
    if (results) *results = s.h_results;
    if (packed_states) *packed_states = s.h_states;
    s.busy = false;
    e->tail = (e->tail + 1) % GACT_MAX_INFLIGHT;
    e->inflight--;
    return rc;
}

int gact_engine_align_tiles(gact_engine *e, int n, const gact_tile_desc *descs,
                            gact_tile_result *results, uint32_t *packed_states)
{
    if (!e) return GACT_ERR_ARG;
    { CU(e, cudaSetDevice(e->device)); int rs = ensure_slots(e); if (rs) return rs; }
    if (e->inflight) return fail(e, GACT_ERR_STATE, "align_tiles while async batches are in flight");
    if (e->staged) return fail(e, GACT_ERR_STATE, "align_tiles while a staged batch is pending");
    CU(e, cudaSetDevice(e->device));
    Slot &s = e->slots[0];
    int rc = enqueue(e, s, n, descs);
    if (rc) return rc;
    if (n > 0) {
        rc = launch_batch(e, s, e->stream, 0);
        if (rc) return rc;
        rc = download(e, s, packed_states != nullptr);
        if (rc) return rc;
    }
    CU(e, cudaEventRecord(s.ev_done, n > 0 ? e->s_d2h : e->stream));
    return finish(e, s, results, packed_states, packed_states != nullptr);
}

int gact_engine_stage(gact_engine *e, int n, const gact_tile_desc *descs)
{
    if (!e) return GACT_ERR_ARG;
    { CU(e, cudaSetDevice(e->device)); int rs = ensure_slots(e); if (rs) return rs; }
    if (e->inflight) return fail(e, GACT_ERR_STATE, "stage while async batches are in flight");
    CU(e, cudaSetDevice(e->device));
    Slot &s = e->slots[0];
    int rc = enqueue(e, s, n, descs);
    if (rc) return rc;
    CU(e, cudaStreamSynchronize(e->s_h2d));
    CU(e, cudaStreamSynchronize(e->stream));
    e->staged = true;
    return GACT_OK;
}

int gact_engine_run_staged(gact_engine *e)
{
    if (!e) return GACT_ERR_ARG;
    if (!e->staged) return fail(e, GACT_ERR_STATE, "run_staged without stage");
    CU(e, cudaSetDevice(e->device));
    Slot &s = e->slots[0];
    if (s.n == 0) return GACT_OK;
    int rc = launch_batch(e, s, e->stream, 0);
    if (rc) return rc;
    e->stats.tiles += s.n;
    e->stats.cells += s.cells;
    e->stats.first_tiles += s.n_first;
    e->stats.batches++;
    return GACT_OK;
}

int gact_engine_sync(gact_engine *e)
{
    if (!e) return GACT_ERR_ARG;
    CU(e, cudaSetDevice(e->device));
    for (int k = 0; k < GACT_MAX_INFLIGHT; k++) CU(e, cudaStreamSynchronize(e->cs[k]));
    CU(e, cudaStreamSynchronize(e->stream));
    return GACT_OK;
}

double gact_engine_last_kernel_ms(gact_engine *e)
{
    if (!e) return -1.0;
    if (e->staged) {
        Slot &s = e->slots[0];
        if (s.n == 0) return -1.0;
        if (cudaSetDevice(e->device) != cudaSuccess) return -1.0;
        if (cudaEventSynchronize(s.ev_k1) != cudaSuccess) return -1.0;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1) != cudaSuccess) { cudaGetLastError(); return -1.0; }
        return ms;
    }
    return e->last_kernel_ms;
}

int gact_engine_fetch_staged(gact_engine *e, gact_tile_result *results, uint32_t *packed_states)
{
    if (!e) return GACT_ERR_ARG;
    if (!e->staged) return fail(e, GACT_ERR_STATE, "fetch_staged without stage");
    CU(e, cudaSetDevice(e->device));
    Slot &s = e->slots[0];
    int rc = download(e, s, packed_states != nullptr);
    if (rc) return rc;
    CU(e, cudaStreamSynchronize(e->s_d2h));
    CU(e, cudaStreamSynchronize(e->stream));
    if (s.n > 0) {
        e->last_n_it = s.n_it;
        e->last_n_escaped = s.n_it > 0 ? s.h_it_info[1] : 0;
        if (results) memcpy(results, s.h_results, (size_t)s.n * sizeof(gact_tile_result));
        if (packed_states) memcpy(packed_states, s.h_states, (size_t)s.n * e->pitch_words * 4);
    }
    e->staged = false;
    return GACT_OK;
}

int gact_engine_tile_path_info(const gact_engine *e, int *n_inter_task, int *n_handed_back)
{
    if (!e) return GACT_ERR_ARG;
    if (n_inter_task) *n_inter_task = e->last_n_it;
    if (n_handed_back) *n_handed_back = e->last_n_escaped;
    return GACT_OK;
}

int gact_engine_reserve_tiles(gact_engine *e)
{
    if (!e) return GACT_ERR_ARG;
    CU(e, cudaSetDevice(e->device));
    return ensure_slots(e);
}

int gact_engine_stats(const gact_engine *e, gact_stats *out)
{
    if (!e || !out) return GACT_ERR_ARG;
    *out = e->stats;
    return GACT_OK;
}
int gact_engine_reset_stats(gact_engine *e)
{
    if (!e) return GACT_ERR_ARG;
    e->stats = gact_stats{};
    return GACT_OK;
}

int gact_engine_set_kernel(gact_engine *e, int variant)
{
    if (!e) return GACT_ERR_ARG;
    if (variant < 0 || variant > 2) return fail(e, GACT_ERR_ARG, "unknown kernel variant");
    if (variant == 2 && !e->s16h.ok) return fail(e, GACT_ERR_ARG, "s16x2 kernel cannot run these parameters");
    e->variant_req = variant;
    return GACT_OK;
}
int gact_engine_get_kernel(const gact_engine *e)
{
    if (!e) return GACT_ERR_ARG;
    return use_s16(e) ? 2 : 1;
}

}  // extern "C"

// ===========================================================================
// D-SOFT on the device
namespace {
// one set of query / candidate buffers: two of them, so that the next batch of reads can be filtered while the
// host still reads the previous batch's candidates
struct DsoftBuf {
    DsoftQuery *d_queries = nullptr, *h_queries = nullptr;
    DsoftCand *d_out = nullptr;
    gact_dsoft_cand *h_out = nullptr;                   // pinned
    unsigned long long *d_count = nullptr, *h_count = nullptr;
    int *d_counter = nullptr;
    size_t q_cap = 0, out_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_done = nullptr, ev_fork = nullptr;
    int n_queries = 0;
    size_t limit = 0;                                   // candidates the kernel may write / the copy brings back
    bool busy = false;
};
constexpr int DSOFT_BUFS = 2;
}  // namespace

struct gact_dsoft {
    gact_engine *e = nullptr;
    DsoftParams p{};
    uint32_t *d_index = nullptr, *d_pos = nullptr;
    bool owns_tables = true;          // false: the tables belong to a gact_seed_table
    uint32_t *d_keys = nullptr, *d_touched = nullptr;
    unsigned long long *d_vals = nullptr;
    cudaStream_t stream = nullptr;    // the filter's own stream: its kernels overlap the chain kernels of other batches
    DsoftBuf buf[DSOFT_BUFS];
    int head = 0, tail = 0, inflight = 0;
    int ctas = 0;
    double last_ms = -1.0;
};

namespace {
void free_dsoft_buf(DsoftBuf &b)
{
    cudaFree(b.d_queries); cudaFree(b.d_out); cudaFree(b.d_count); cudaFree(b.d_counter);
    if (b.h_queries) cudaFreeHost(b.h_queries);
    if (b.h_out) cudaFreeHost(b.h_out);
    if (b.h_count) cudaFreeHost(b.h_count);
    for (cudaEvent_t ev : {b.ev0, b.ev1, b.ev_done, b.ev_fork}) if (ev) cudaEventDestroy(ev);
    b = DsoftBuf();
}

int reserve_dsoft_buf(gact_engine *e, DsoftBuf &b, size_t n_queries, size_t out_cap)
{
    if (!b.ev0) {
        CU(e, cudaEventCreate(&b.ev0));
        CU(e, cudaEventCreate(&b.ev1));
        CU(e, cudaEventCreateWithFlags(&b.ev_done, cudaEventDisableTiming));
        CU(e, cudaEventCreateWithFlags(&b.ev_fork, cudaEventDisableTiming));
    }
    if (!b.d_count) {
        if (cudaMalloc(&b.d_count, 8) != cudaSuccess || cudaMalloc(&b.d_counter, 4) != cudaSuccess ||
            cudaMallocHost(&b.h_count, 8) != cudaSuccess) { cudaGetLastError(); return fail(e, GACT_ERR_NOMEM, "cudaMalloc(D-SOFT counters) failed"); }
    }
    if (n_queries > b.q_cap) {
        cudaFree(b.d_queries); if (b.h_queries) cudaFreeHost(b.h_queries);
        b.d_queries = nullptr; b.h_queries = nullptr; b.q_cap = 0;
        const size_t want = std::max<size_t>(n_queries, 64);
        if (cudaMalloc(&b.d_queries, want * sizeof(DsoftQuery)) != cudaSuccess ||
            cudaMallocHost(&b.h_queries, want * sizeof(DsoftQuery)) != cudaSuccess) { cudaGetLastError(); return fail(e, GACT_ERR_NOMEM, "cudaMalloc(queries) failed"); }
        b.q_cap = want;
    }
    if (out_cap > b.out_cap || !b.d_out) {
        cudaFree(b.d_out); if (b.h_out) cudaFreeHost(b.h_out);
        b.d_out = nullptr; b.h_out = nullptr; b.out_cap = 0;
        const size_t want = std::max<size_t>(out_cap, 1024);
        if (cudaMalloc(&b.d_out, want * sizeof(DsoftCand)) != cudaSuccess ||
            cudaMallocHost(&b.h_out, want * sizeof(gact_dsoft_cand)) != cudaSuccess) { cudaGetLastError(); return fail(e, GACT_ERR_NOMEM, "cudaMalloc(candidates) failed"); }
        b.out_cap = want;
    }
    return GACT_OK;
}
}  // namespace

extern "C" {

void gact_dsoft_destroy(gact_dsoft *d)
{
    if (!d) return;
    cudaSetDevice(d->e->device);
    cudaStreamSynchronize(d->e->stream);
    if (d->stream) cudaStreamSynchronize(d->stream);
    if (d->owns_tables) { cudaFree(d->d_index); cudaFree(d->d_pos); }
    cudaFree(d->d_keys); cudaFree(d->d_touched); cudaFree(d->d_vals);
    for (int k = 0; k < DSOFT_BUFS; k++) free_dsoft_buf(d->buf[k]);
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d;
}

// everything of a D-SOFT handle except the seed-position table itself
static int dsoft_alloc(gact_dsoft **out, gact_engine *e, int kmer_size, int window_size, uint32_t bin_size,
                       uint32_t kmer_max_occurence, int num_seeds, int threshold, int max_candidates)
{
    if (kmer_size < 4 || kmer_size > 15 || window_size < 1 || window_size > 32 || window_size >= kmer_size ||
        bin_size == 0 || num_seeds < 0 || threshold < 1)
        return fail(e, GACT_ERR_ARG, "bad D-SOFT parameters");
    CU(e, cudaSetDevice(e->device));
    gact_dsoft *d = new (std::nothrow) gact_dsoft();
    if (!d) return fail(e, GACT_ERR_NOMEM, "host allocation failed");
    d->e = e;
    d->owns_tables = false;
    // every used seed can touch at most max_occ bins: size the per-warp table for the worst case, load <= 0.5
    uint64_t need = 2ull * ((uint64_t)num_seeds + 2) * std::max<uint32_t>(kmer_max_occurence, 1u);
    uint32_t cap = 1024;
    while (cap < need && cap < (1u << 24)) cap <<= 1;
    if (cap < need) { delete d; return fail(e, GACT_ERR_ARG, "D-SOFT table would exceed 16M slots per warp"); }
    // The filter is bound by dependent memory latency (bucket -> hits -> bin slot, seed after seed), so it wants
    // as many strands in flight as fit: up to 32 warps per SM while the per-warp tables stay under 8 GiB,
    // never fewer than 8 warps per SM.
    int ctas_per_sm = 8;
    while (ctas_per_sm > 2 && (uint64_t)e->num_sms * ctas_per_sm * 4 * cap * 16ull > (8ull << 30)) ctas_per_sm--;
    d->ctas = e->num_sms * ctas_per_sm;
    const size_t warps = (size_t)d->ctas * 4;
    bool ok = cudaMalloc(&d->d_keys, warps * cap * 4) == cudaSuccess &&
              cudaMalloc(&d->d_touched, warps * cap * 4) == cudaSuccess &&
              cudaMalloc(&d->d_vals, warps * cap * 8) == cudaSuccess &&
              cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) { cudaGetLastError(); gact_dsoft_destroy(d); return fail(e, GACT_ERR_NOMEM, "cudaMalloc(D-SOFT tables) failed"); }
    cudaMemsetAsync(d->d_keys, 0, warps * cap * 4, d->stream);
    d->p.k = kmer_size; d->p.w = window_size; d->p.bin_size = bin_size; d->p.max_occ = kmer_max_occurence;
    d->p.num_seeds = num_seeds; d->p.threshold = threshold; d->p.max_candidates = max_candidates;
    d->p.table_cap = cap;
    *out = d;
    return GACT_OK;
}

int gact_dsoft_create(gact_dsoft **out, gact_engine *e, const uint32_t *index_table, uint64_t index_entries,
                      const uint32_t *pos_table, uint64_t n_pos, int kmer_size, int window_size,
                      uint32_t bin_size, uint32_t kmer_max_occurence, int num_seeds, int threshold, int max_candidates)
{
    if (!out || !e || !index_table || (!pos_table && n_pos)) return GACT_ERR_ARG;
    *out = nullptr;
    if (kmer_size < 4 || kmer_size > 15 || index_entries != ((uint64_t)1 << (2 * kmer_size)) + 1)
        return fail(e, GACT_ERR_ARG, "bad D-SOFT parameters");
    gact_dsoft *d = nullptr;
    int rc = dsoft_alloc(&d, e, kmer_size, window_size, bin_size, kmer_max_occurence, num_seeds, threshold, max_candidates);
    if (rc) return rc;
    d->owns_tables = true;
    if (cudaMalloc(&d->d_index, index_entries * 4) != cudaSuccess ||
        cudaMalloc(&d->d_pos, std::max<uint64_t>(n_pos, 1) * 4) != cudaSuccess) {
        cudaGetLastError(); gact_dsoft_destroy(d);
        return fail(e, GACT_ERR_NOMEM, "cudaMalloc(D-SOFT tables) failed");
    }
    cudaMemcpyAsync(d->d_index, index_table, index_entries * 4, cudaMemcpyHostToDevice, d->stream);
    if (n_pos) cudaMemcpyAsync(d->d_pos, pos_table, n_pos * 4, cudaMemcpyHostToDevice, d->stream);
    cudaError_t r = cudaStreamSynchronize(d->stream);
    if (r != cudaSuccess) { gact_dsoft_destroy(d); return fail(e, GACT_ERR_CUDA, std::string("D-SOFT upload: ") + cudaGetErrorString(r)); }
    e->stats.h2d_bytes += (double)(index_entries + n_pos) * 4;
    d->p.index_table = d->d_index; d->p.pos_table = d->d_pos;
    *out = d;
    return GACT_OK;
}

// ---- seed-position table built on the device (seed_build.cuh) -------------------------------
struct gact_seed_table {
    gact_engine *e = nullptr;
    SeedTableDev t;
};

int gact_seed_table_build(gact_seed_table **out, gact_engine *e, const char *ref, uint32_t ref_len, int kmer_size,
                          uint32_t seed_occurence_multiple, uint32_t bin_size, uint32_t window_size)
{
    if (!out || !e || (!ref && ref_len)) return GACT_ERR_ARG;
    *out = nullptr;
    if (e->inflight || e->staged) return fail(e, GACT_ERR_STATE, "seed table build while batches are outstanding");
    CU(e, cudaSetDevice(e->device));
    gact_seed_table *t = new (std::nothrow) gact_seed_table();
    if (!t) return fail(e, GACT_ERR_NOMEM, "host allocation failed");
    t->e = e;
    std::string err;
    const int rc = seed_table_build_device(ref ? ref : "", ref_len, kmer_size, window_size, seed_occurence_multiple, bin_size,
                                           e->stream, &t->t, &err);
    if (rc) {
        delete t;
        return fail(e, rc == 1 ? GACT_ERR_ARG : rc == 2 ? GACT_ERR_NOMEM : GACT_ERR_CUDA, "seed table: " + err);
    }
    e->stats.h2d_bytes += (double)ref_len;
    *out = t;
    return GACT_OK;
}

void gact_seed_table_destroy(gact_seed_table *t)
{
    if (!t) return;
    cudaSetDevice(t->e->device);
    cudaStreamSynchronize(t->e->stream);
    seed_table_free(&t->t);
    delete t;
}

int gact_seed_table_info(const gact_seed_table *t, uint64_t *index_entries, uint32_t *num_minimizers,
                         uint32_t *kmer_max_occurence, double *build_ms)
{
    if (!t) return GACT_ERR_ARG;
    if (index_entries) *index_entries = t->t.index_entries;
    if (num_minimizers) *num_minimizers = t->t.n_pos;
    if (kmer_max_occurence) *kmer_max_occurence = t->t.max_occ;
    if (build_ms) *build_ms = t->t.build_ms;
    return GACT_OK;
}

int gact_seed_table_download(const gact_seed_table *t, uint32_t *index_table, uint32_t *pos_table)
{
    if (!t) return GACT_ERR_ARG;
    gact_engine *e = t->e;
    CU(e, cudaSetDevice(e->device));
    if (index_table) CU(e, cudaMemcpy(index_table, t->t.d_index, t->t.index_entries * 4, cudaMemcpyDeviceToHost));
    if (pos_table && t->t.n_pos) CU(e, cudaMemcpy(pos_table, t->t.d_pos, (size_t)t->t.n_pos * 4, cudaMemcpyDeviceToHost));
    return GACT_OK;
}

int gact_dsoft_create_from_table(gact_dsoft **out, gact_engine *e, const gact_seed_table *t, int num_seeds, int threshold,
                                 int max_candidates)
{
    if (!out || !e || !t) return GACT_ERR_ARG;
    *out = nullptr;
    if (t->e->device != e->device) return fail(e, GACT_ERR_ARG, "seed table lives on another device");
    gact_dsoft *d = nullptr;
    int rc = dsoft_alloc(&d, e, t->t.k, t->t.w, t->t.bin_size, t->t.max_occ, num_seeds, threshold, max_candidates);
    if (rc) return rc;
    d->d_index = t->t.d_index; d->d_pos = t->t.d_pos;      // borrowed: the table must outlive the filter
    d->p.index_table = d->d_index; d->p.pos_table = d->d_pos;
    CU(e, cudaStreamSynchronize(e->stream));
    CU(e, cudaStreamSynchronize(d->stream));
    *out = d;
    return GACT_OK;
}

int gact_dsoft_reserve(gact_dsoft *d, int n_queries, int64_t out_cap)
{
    if (!d || n_queries < 0 || out_cap < 0) return GACT_ERR_ARG;
    gact_engine *e = d->e;
    CU(e, cudaSetDevice(e->device));
    for (int k = 0; k < DSOFT_BUFS; k++) {
        if (d->buf[k].busy) continue;
        int rc = reserve_dsoft_buf(e, d->buf[k], (size_t)n_queries, (size_t)out_cap);
        if (rc) return rc;
    }
    return GACT_OK;
}

int gact_dsoft_submit(gact_dsoft *d, int n_queries, const int32_t *sets, const int64_t *seq_index, int64_t out_cap)
{
    if (!d || n_queries < 0 || (n_queries && (!sets || !seq_index)) || out_cap < 0) return GACT_ERR_ARG;
    gact_engine *e = d->e;
    if (d->inflight >= DSOFT_BUFS) return fail(e, GACT_ERR_STATE, "two D-SOFT batches already in flight");
    CU(e, cudaSetDevice(e->device));
    DsoftBuf &b = d->buf[d->head];
    {
        int rr = reserve_dsoft_buf(e, b, (size_t)n_queries, (size_t)out_cap);
        if (rr) return rr;
    }
    for (int i = 0; i < n_queries; i++) {
        const int s = sets[i];
        if (s < 0 || s >= GACT_MAX_SETS) return fail(e, GACT_ERR_ARG, "bad set in D-SOFT query");
        const SeqSetHost &hs = e->sets[s];
        if (seq_index[i] < 0 || (size_t)seq_index[i] + 1 >= hs.starts.size()) return fail(e, GACT_ERR_ARG, "bad sequence index in D-SOFT query");
        b.h_queries[i].start = hs.starts[(size_t)seq_index[i]];
        b.h_queries[i].len = (int)(hs.starts[(size_t)seq_index[i] + 1] - hs.starts[(size_t)seq_index[i]]);
        b.h_queries[i].set = s;
    }
    b.n_queries = n_queries;
    b.limit = std::min<size_t>((size_t)out_cap, b.out_cap);
    *b.h_count = 0;
    cudaStream_t st = d->stream;
    CU(e, cudaEventRecord(b.ev_fork, e->stream));          // ordered after uploads / table builds on the engine's stream
    CU(e, cudaStreamWaitEvent(st, b.ev_fork, 0));
    if (n_queries > 0) {
        for (int i = 0; i < GACT_MAX_SETS; i++) d->p.sets[i] = e->kp.sets[i];
        CU(e, cudaMemcpyAsync(b.d_queries, b.h_queries, (size_t)n_queries * sizeof(DsoftQuery), cudaMemcpyHostToDevice, st));
        CU(e, cudaMemsetAsync(b.d_count, 0, 8, st));
        CU(e, cudaMemsetAsync(b.d_counter, 0, 4, st));
        CU(e, cudaEventRecord(b.ev0, st));
        int ctas = d->ctas;
        if ((n_queries + 3) / 4 < ctas) ctas = (n_queries + 3) / 4;
        // the table area was sized for d->ctas CTAs of 4 warps; fewer CTAs use a prefix of it
        dsoft_kernel<<<ctas, 128, 0, st>>>(d->p, b.d_queries, n_queries, d->d_keys, d->d_vals, d->d_touched,
                                           b.d_out, (unsigned long long)b.limit, b.d_count, b.d_counter);
        CU(e, cudaGetLastError());
        CU(e, cudaEventRecord(b.ev1, st));
        CU(e, cudaMemcpyAsync(b.h_count, b.d_count, 8, cudaMemcpyDeviceToHost, st));
        // the count is not known to the host yet: bring back the whole window the kernel was allowed to fill
        // (16 B per candidate, a few candidates per read)
        if (b.limit) CU(e, cudaMemcpyAsync(b.h_out, b.d_out, b.limit * sizeof(DsoftCand), cudaMemcpyDeviceToHost, st));
        e->stats.kernel_launches++;
    }
    CU(e, cudaEventRecord(b.ev_done, st));
    b.busy = true;
    d->head = (d->head + 1) % DSOFT_BUFS;
    d->inflight++;
    return GACT_OK;
}

int gact_dsoft_wait(gact_dsoft *d, gact_dsoft_cand *out, int64_t out_cap, int64_t *n_out)
{
    if (!d || !n_out || out_cap < 0 || (out_cap && !out)) return GACT_ERR_ARG;
    gact_engine *e = d->e;
    *n_out = 0;
    if (d->inflight == 0) return fail(e, GACT_ERR_STATE, "dsoft_wait without dsoft_submit");
    CU(e, cudaSetDevice(e->device));
    DsoftBuf &b = d->buf[d->tail];
    CU(e, cudaEventSynchronize(b.ev_done));
    b.busy = false;
    d->tail = (d->tail + 1) % DSOFT_BUFS;
    d->inflight--;
    if (b.n_queries == 0) return GACT_OK;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, b.ev0, b.ev1);
    d->last_ms = ms;
    const unsigned long long total = *b.h_count;
    *n_out = (int64_t)total;
    if (total > b.limit || (int64_t)total > out_cap) return fail(e, GACT_ERR_NOMEM, "candidate buffer too small (see *n_out)");
    if (total) {
        static_assert(sizeof(DsoftCand) == sizeof(gact_dsoft_cand), "candidate layouts differ");
        // group by query, emission order inside a query: every candidate carries its query and its index inside that
        // query, so its final position is known without a comparison sort
        std::vector<uint32_t> start((size_t)b.n_queries + 1, 0);
        bool ok = true;
        for (unsigned long long i = 0; i < total && ok; i++) {
            const int32_t q = b.h_out[i].query;
            if (q < 0 || q >= b.n_queries) ok = false; else start[(size_t)q + 1]++;
        }
        for (int q = 0; q < b.n_queries; q++) start[(size_t)q + 1] += start[(size_t)q];
        for (unsigned long long i = 0; i < total && ok; i++) {
            const gact_dsoft_cand &c = b.h_out[i];
            const uint32_t pos = start[(size_t)c.query] + (uint32_t)c.seq;
            if (c.seq < 0 || pos >= start[(size_t)c.query + 1]) ok = false; else out[pos] = c;
        }
        if (!ok) {                                         // cannot happen with a well-formed candidate stream
            memcpy(out, b.h_out, (size_t)total * sizeof(gact_dsoft_cand));
            std::sort(out, out + total, [](const gact_dsoft_cand &a, const gact_dsoft_cand &c) {
                return a.query != c.query ? a.query < c.query : a.seq < c.seq;
            });
        }
        e->stats.d2h_bytes += (double)total * sizeof(DsoftCand);
    }
    return GACT_OK;
}

int gact_dsoft_run(gact_dsoft *d, int n_queries, const int32_t *sets, const int64_t *seq_index,
                   gact_dsoft_cand *out, int64_t out_cap, int64_t *n_out)
{
    if (!d || n_queries < 0 || (n_queries && (!sets || !seq_index)) || !n_out || out_cap < 0 || (out_cap && !out))
        return GACT_ERR_ARG;
    *n_out = 0;
    if (d->inflight) return fail(d->e, GACT_ERR_STATE, "dsoft_run while asynchronous D-SOFT batches are in flight");
    if (n_queries == 0) return GACT_OK;
    int rc = gact_dsoft_submit(d, n_queries, sets, seq_index, out_cap);
    if (rc) return rc;
    return gact_dsoft_wait(d, out, out_cap, n_out);
}

double gact_dsoft_last_kernel_ms(const gact_dsoft *d) { return d ? d->last_ms : -1.0; }

}  // extern "C"

// ===========================================================================
// whole candidate extensions on the device
namespace {

// resources of one extend batch in flight
struct ChainBatch {
    ChainCall *d_calls = nullptr, *h_calls = nullptr;       // device / pinned host
    ChainResult *d_res = nullptr, *h_res = nullptr;
    ChainAux *d_aux = nullptr;                              // [2]: main launch, long-chain lane
    int *d_taken = nullptr;
    size_t cap = 0;
    std::vector<int> perm;                                  // device order -> caller's order
    int n = 0;
    bool busy = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_lane = nullptr, ev_done = nullptr;
};

void free_chain_batch(ChainBatch &b)
{
    if (b.d_calls) cudaFree(b.d_calls);
    if (b.d_res) cudaFree(b.d_res);
    if (b.d_aux) cudaFree(b.d_aux);
    if (b.d_taken) cudaFree(b.d_taken);
    if (b.h_calls) cudaFreeHost(b.h_calls);
    if (b.h_res) cudaFreeHost(b.h_res);
    for (cudaEvent_t ev : {b.ev0, b.ev1, b.ev_fork, b.ev_lane, b.ev_done}) if (ev) cudaEventDestroy(ev);
    b = ChainBatch();
}

int reserve_chain_batch(gact_engine *e, ChainBatch &b, size_t n)
{
    if (!b.ev0) {
        CU(e, cudaEventCreate(&b.ev0));
        CU(e, cudaEventCreate(&b.ev1));
        CU(e, cudaEventCreateWithFlags(&b.ev_fork, cudaEventDisableTiming));
        CU(e, cudaEventCreateWithFlags(&b.ev_lane, cudaEventDisableTiming));
        CU(e, cudaEventCreateWithFlags(&b.ev_done, cudaEventDisableTiming));
    }
    if (n <= b.cap) return GACT_OK;
    if (b.d_calls) cudaFree(b.d_calls);
    if (b.d_res) cudaFree(b.d_res);
    if (b.d_taken) cudaFree(b.d_taken);
    if (b.h_calls) cudaFreeHost(b.h_calls);
    if (b.h_res) cudaFreeHost(b.h_res);
    b.d_calls = nullptr; b.d_res = nullptr; b.d_taken = nullptr; b.h_calls = nullptr; b.h_res = nullptr; b.cap = 0;
    const size_t want = std::max<size_t>(n, 256);
    if ((!b.d_aux && cudaMalloc(&b.d_aux, 2 * sizeof(ChainAux)) != cudaSuccess) ||
        cudaMalloc(&b.d_calls, want * sizeof(ChainCall)) != cudaSuccess ||
        cudaMalloc(&b.d_res, want * sizeof(ChainResult)) != cudaSuccess ||
        cudaMalloc(&b.d_taken, want * sizeof(int)) != cudaSuccess ||
        cudaMallocHost(&b.h_calls, want * sizeof(ChainCall)) != cudaSuccess ||
        cudaMallocHost(&b.h_res, want * sizeof(ChainResult)) != cudaSuccess) {
        cudaGetLastError();
        return fail(e, GACT_ERR_NOMEM, "allocation of the chain buffers failed");
    }
    b.cap = want;
    return GACT_OK;
}

}  // namespace

struct gact_chain_state {
    ChainBatch batch[GACT_MAX_INFLIGHT];
    cudaStream_t lane[GACT_MAX_INFLIGHT] = {};       // second stream of a batch: the long-chain lane
    int head = 0, tail = 0, inflight = 0;
    int mode = 0;                                    // gact_engine_set_chain_mode
    int last_mode = 0, last_ctas = 0, last_long = 0; // what the last submit chose (diagnostics)
};

static void destroy_chain_state(gact_engine *e)
{
    if (!e->chains) return;
    for (int k = 0; k < GACT_MAX_INFLIGHT; k++) {
        if (e->chains->lane[k]) { cudaStreamSynchronize(e->chains->lane[k]); cudaStreamDestroy(e->chains->lane[k]); }
        free_chain_batch(e->chains->batch[k]);
    }
    delete e->chains;
    e->chains = nullptr;
}

extern "C" {

int gact_engine_extend_supported(const gact_engine *e)
{
    // exceptions (bytes other than ACGT) in the reference are handled inside the chain kernels; a QUERY sequence that holds
    // one must go through the tile path (gact_engine_seq_has_exceptions tells which)
    if (!e || !e->s16h.ok || !e->s16h.lut_ok || e->variant_req == 1) return 0;
    return 1;
}

static gact_chain_state *chain_state(gact_engine *e)
{
    if (!e->chains) e->chains = new (std::nothrow) gact_chain_state();
    return e->chains;
}

int gact_engine_extend_reserve(gact_engine *e, int n)
{
    if (!e || n < 0) return GACT_ERR_ARG;
    CU(e, cudaSetDevice(e->device));
    gact_chain_state *cs = chain_state(e);
    if (!cs) return fail(e, GACT_ERR_NOMEM, "host allocation failed");
    for (int k = 0; k < GACT_MAX_INFLIGHT; k++) {
        if (cs->batch[k].busy) continue;
        int rc = reserve_chain_batch(e, cs->batch[k], (size_t)n);
        if (rc) return rc;
        if (!cs->lane[k]) CU(e, cudaStreamCreateWithFlags(&cs->lane[k], cudaStreamNonBlocking));
    }
    return GACT_OK;
}

int gact_engine_set_chain_mode(gact_engine *e, int mode)
{
    if (!e || mode < 0 || mode > 4) return GACT_ERR_ARG;
    gact_chain_state *cs = chain_state(e);
    if (!cs) return fail(e, GACT_ERR_NOMEM, "host allocation failed");
    if ((mode == 1 || mode == 2 || mode == 4) && !e->s16h_lat.ok)
        return fail(e, GACT_ERR_ARG, "the latency chain kernel needs tile_size <= 320 and the one-PRMT score table");
    cs->mode = mode;
    return GACT_OK;
}

int gact_engine_extend_submit(gact_engine *e, int n, const gact_call *calls)
{
    if (!e || n < 0 || (n > 0 && !calls)) return GACT_ERR_ARG;
    if (!gact_engine_extend_supported(e))
        return fail(e, GACT_ERR_ARG, "on-device extension needs scores in the packed kernels' 16-bit range and ACGT-only sets");
    if (e->inflight || e->staged) return fail(e, GACT_ERR_STATE, "extend while tile batches are outstanding");
    gact_chain_state *cs = chain_state(e);
    if (!cs) return fail(e, GACT_ERR_NOMEM, "host allocation failed");
    if (cs->inflight >= GACT_MAX_INFLIGHT) return fail(e, GACT_ERR_STATE, "GACT_MAX_INFLIGHT extend batches already in flight");
    CU(e, cudaSetDevice(e->device));
    const int slot = cs->head;
    ChainBatch &b = cs->batch[slot];
    {
        int rr = reserve_chain_batch(e, b, (size_t)n);
        if (rr) return rr;
        if (!cs->lane[slot]) CU(e, cudaStreamCreateWithFlags(&cs->lane[slot], cudaStreamNonBlocking));
    }
    b.n = n;
    cudaStream_t st = e->cs[slot];
    CU(e, cudaEventRecord(b.ev_fork, e->stream));           // ordered after what the caller enqueued on the engine's stream
    CU(e, cudaStreamWaitEvent(st, b.ev_fork, 0));
    if (n > 0) {
        const SeqSetHost &rs = e->sets[GACT_SET_REF];
        const int et = std::max(1, e->params.tile_size - e->params.tile_overlap);
        // longest query first: a chain is a serial run of tiles, so the long ones must start early or they
        // are still running alone when every other chain slot has drained.  Counting sort on the expected tile count.
        b.perm.resize((size_t)n);
        std::vector<int> est((size_t)n);
        int est_max = 0;
        double est_sum = 0.0;
        for (int i = 0; i < n; i++) {
            const gact_call &c = calls[i];
            if (c.query_set >= GACT_MAX_SETS || c.ref_seq < 0 || (size_t)c.ref_seq + 1 >= rs.starts.size())
                return fail(e, GACT_ERR_ARG, "call " + std::to_string(i) + " out of range");
            const SeqSetHost &qs = e->sets[c.query_set];
            if (c.query_seq < 0 || (size_t)c.query_seq + 1 >= qs.starts.size())
                return fail(e, GACT_ERR_ARG, "call " + std::to_string(i) + " out of range");
            const long long ql = qs.starts[(size_t)c.query_seq + 1] - qs.starts[(size_t)c.query_seq];
            const long long rl = rs.starts[(size_t)c.ref_seq + 1] - rs.starts[(size_t)c.ref_seq];
            if (c.ref_pos < 0 || c.query_pos < 0 || c.ref_pos > rl || c.query_pos > ql)
                return fail(e, GACT_ERR_ARG, "call " + std::to_string(i) + ": anchor outside its sequences");
            if (!qs.seq_exc.empty() && qs.seq_exc[(size_t)c.query_seq])
                return fail(e, GACT_ERR_ARG, "call " + std::to_string(i) + ": the query sequence holds bytes other than ACGT "
                                             "(gact_engine_seq_has_exceptions); extend it through the tile path");
            est[(size_t)i] = (int)std::min<long long>(std::min(ql, rl) / et + 2, 1 << 20);
            est_max = std::max(est_max, est[(size_t)i]);
            est_sum += est[(size_t)i];
        }
        if (getenv("GACT_CHAIN_NOSORT")) {
            for (int i = 0; i < n; i++) b.perm[(size_t)i] = i;
        } else {
            std::vector<int> start((size_t)est_max + 2, 0);
            for (int i = 0; i < n; i++) start[(size_t)(est_max - est[(size_t)i]) + 1]++;
            for (int k = 1; k <= est_max + 1; k++) start[(size_t)k] += start[(size_t)k - 1];
            for (int i = 0; i < n; i++) b.perm[(size_t)start[(size_t)(est_max - est[(size_t)i])]++] = i;
        }
        for (int i = 0; i < n; i++) {
            const gact_call &c = calls[b.perm[(size_t)i]];
            const SeqSetHost &qs = e->sets[c.query_set];
            ChainCall &d = b.h_calls[i];
            d.ref_start = rs.starts[(size_t)c.ref_seq];
            d.query_start = qs.starts[(size_t)c.query_seq];
            d.ref_len = (int)(rs.starts[(size_t)c.ref_seq + 1] - d.ref_start);
            d.query_len = (int)(qs.starts[(size_t)c.query_seq + 1] - d.query_start);
            d.ref_pos = c.ref_pos; d.query_pos = c.query_pos;
            d.query_set = c.query_set; d.pad = 0;
        }
        CU(e, cudaMemcpyAsync(b.d_calls, b.h_calls, (size_t)n * sizeof(ChainCall), cudaMemcpyHostToDevice, st));
        CU(e, cudaEventRecord(b.ev0, st));

        // ---- which mapping ----
        //   1: latency kernel, one CTA (4 warps = one per SM sub-partition) per SM    2: latency kernel, two CTAs per SM
        //   3: throughput kernel (two tiles per warp for tile_size <= 320)
        //   4: split -- the longest chains on the latency kernel (one CTA per SM, own stream), the rest on the throughput kernel
        // auto: about half of the candidates die in their first tile, so the real load per sub-partition is ~ est_sum / 2 / (4 SMs);
        // while that stays below the longest chain, the longest chain sets the time and must run alone on its sub-partition.
        const bool lat_ok = e->s16h_lat.ok;
        int mode = cs->mode;
        if (const char *m = getenv("GACT_CHAIN_MODE")) mode = atoi(m);
        const double per_sub = 0.5 * est_sum / (4.0 * e->num_sms);
        const double thr_slots = (double)e->s16h.slots();
        if (mode == 0) {
            // measured on B200 (profiles/r2_chain_modes.txt; 1/8, 1/4, 1/2 and the whole of a 50 MB read set):
            // load per sub-partition below 0.6 longest chains -> every long chain alone on a sub-partition;
            // up to 2 longest chains -> the 4 * SMs longest chains on the latency kernel beside the throughput kernel; above
            // that the shard is work-bound and the throughput kernel alone is fastest
            if (!lat_ok) mode = 3;
            else if (per_sub <= 0.6 * est_max) mode = 1;
            else if (per_sub <= 2.0 * est_max) mode = 4;
            else mode = 3;
        }
        if (!lat_ok && mode != 3) mode = 3;
        int n_long = 0;
        if (mode == 4) {
            // the longest chains, one per SM sub-partition at most, as long as their serial length exceeds 60 % of the
            // average load per chain slot of the throughput kernel
            const double long_tiles = 0.6 * est_sum / thr_slots;
            const int cap_long = 4 * e->num_sms;
            while (n_long < n && n_long < cap_long && (double)est[(size_t)b.perm[(size_t)n_long]] > long_tiles) n_long++;
            if (n_long < 8) { n_long = 0; mode = 3; }
        }
        cs->last_mode = mode; cs->last_long = n_long;
        const int thr = e->params.first_tile_score_threshold;
        if (mode == 1 || mode == 2) {
            int ctas = (mode == 1 ? 1 : 2) * e->num_sms;
            if (const char *c = getenv("GACT_CHAIN_CTAS")) ctas = atoi(c);
            cs->last_ctas = ctas;
            s16h_launch_chain(e->s16h_lat, e->kp, b.d_calls, n, b.d_res, thr, b.d_aux, b.d_taken, st, 2, ctas, e->num_sms);
            e->stats.kernel_launches++;
        } else {
            if (n_long > 0) {
                CU(e, cudaEventRecord(b.ev_lane, st));
                CU(e, cudaStreamWaitEvent(cs->lane[slot], b.ev_lane, 0));
                s16h_launch_chain(e->s16h_lat, e->kp, b.d_calls, n_long, b.d_res, thr, b.d_aux + 1, b.d_taken, cs->lane[slot], 2,
                                  e->num_sms, e->num_sms);
                CU(e, cudaGetLastError());
                CU(e, cudaEventRecord(b.ev_lane, cs->lane[slot]));
                e->stats.kernel_launches++;
            }
            const int n_rest = n - n_long;
            if (n_rest > 0) {
                // every chain resident: deal the (sorted) chains round-robin over the CTAs so that the long ones do not
                // share SMs; with more chains than slots the plain in-order claim measured faster (profiles/r1_chain_order.txt)
                int deal = (n_rest <= e->s16h.slots()) ? 1 : 0;
                if (const char *d = getenv("GACT_CHAIN_DEAL")) deal = atoi(d);
                cs->last_ctas = e->s16h.ctas;
                s16h_launch_chain(e->s16h, e->kp, b.d_calls + n_long, n_rest, b.d_res + n_long, thr, b.d_aux, b.d_taken + n_long, st,
                                  deal, e->s16h.ctas, e->num_sms, slot);
                e->stats.kernel_launches++;
            }
            if (n_long > 0) CU(e, cudaStreamWaitEvent(st, b.ev_lane, 0));
        }
        CU(e, cudaGetLastError());
        CU(e, cudaEventRecord(b.ev1, st));
        CU(e, cudaMemcpyAsync(b.h_res, b.d_res, (size_t)n * sizeof(ChainResult), cudaMemcpyDeviceToHost, st));
        e->stats.h2d_bytes += (double)n * sizeof(ChainCall);
        e->stats.d2h_bytes += (double)n * sizeof(ChainResult);
    }
    CU(e, cudaEventRecord(b.ev_done, st));
    b.busy = true;
    cs->head = (cs->head + 1) % GACT_MAX_INFLIGHT;
    cs->inflight++;
    return GACT_OK;
}

int gact_engine_extend_wait(gact_engine *e, gact_alignment *out)
{
    if (!e) return GACT_ERR_ARG;
    gact_chain_state *cs = e->chains;
    if (!cs || cs->inflight == 0) return fail(e, GACT_ERR_STATE, "extend_wait without extend_submit");
    CU(e, cudaSetDevice(e->device));
    ChainBatch &b = cs->batch[cs->tail];
    if (b.n > 0 && !out) return GACT_ERR_ARG;
    CU(e, cudaEventSynchronize(b.ev_done));
    if (b.n > 0) {
        float ms = 0.f;
        CU(e, cudaEventElapsedTime(&ms, b.ev0, b.ev1));
        e->last_kernel_ms = ms;
        e->stats.kernel_ms += ms;
        for (int i = 0; i < b.n; i++) {
            const ChainResult &r = b.h_res[i];
            gact_alignment &o = out[b.perm[(size_t)i]];
            o.ab = r.ab; o.ae = r.ae; o.bb = r.bb; o.be = r.be; o.score = r.score; o.first_tile_score = r.first_tile_score;
            o.n_tiles = r.n_tiles; o.reserved = 0; o.n_cells = r.n_cells;
            e->stats.tiles += (uint64_t)r.n_tiles;
            e->stats.cells += (uint64_t)r.n_cells;
        }
    }
    e->stats.batches++;
    b.busy = false;
    cs->tail = (cs->tail + 1) % GACT_MAX_INFLIGHT;
    cs->inflight--;
    return GACT_OK;
}

int gact_engine_extend(gact_engine *e, int n, const gact_call *calls, gact_alignment *out)
{
    if (!e || n < 0 || (n > 0 && (!calls || !out))) return GACT_ERR_ARG;
    if (e->chains && e->chains->inflight) return fail(e, GACT_ERR_STATE, "extend while asynchronous extend batches are in flight");
    int rc = gact_engine_extend_submit(e, n, calls);
    if (rc) return rc;
    return gact_engine_extend_wait(e, out);
}

int gact_engine_chain_info(const gact_engine *e, int *mode, int *ctas, int *n_long)
{
    if (!e || !e->chains) return GACT_ERR_ARG;
    if (mode) *mode = e->chains->last_mode;
    if (ctas) *ctas = e->chains->last_ctas;
    if (n_long) *n_long = e->chains->last_long;
    return GACT_OK;
}

}  // extern "C"

#ifdef GACT_PROF
// profiling build only (libgact_b200_prof.so): read and clear the chain kernel's phase clocks
extern "C" int gact_prof_read(unsigned long long *out8)
{
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(out8, gact::g_chain_prof, sizeof(z)) != cudaSuccess) return GACT_ERR_CUDA;
    if (cudaMemcpyToSymbol(gact::g_chain_prof, z, sizeof(z)) != cudaSuccess) return GACT_ERR_CUDA;
    return GACT_OK;
}
#endif

#ifdef GACT_CHECK
// bounds-checked build only (libgact_b200_check.so): read and clear the violation counters
extern "C" int gact_check_read(unsigned long long *out8)
{
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaDeviceSynchronize() != cudaSuccess) return GACT_ERR_CUDA;
    if (cudaMemcpyFromSymbol(out8, gact::g_check_fail, sizeof(z)) != cudaSuccess) return GACT_ERR_CUDA;
    if (cudaMemcpyToSymbol(gact::g_check_fail, z, sizeof(z)) != cudaSuccess) return GACT_ERR_CUDA;
    return GACT_OK;
}
#endif
