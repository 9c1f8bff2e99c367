#include "gact_scheduler.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace darwin {

// Per-candidate state (the locals of GACT(), gact.cpp:51-80).
struct GactScheduler::Active {
    int32_t call;                 // index into calls / out
    int32_t ref_pos, query_pos;   // moving anchor
    int32_t rev_ref_pos, rev_query_pos;   // where the right extension starts (gact.cpp:72-73,104-105)
    int32_t ab, bb;
    int32_t score;                // running total of gact.cpp:197-210
    int32_t first_tile_score;
    int32_t n_tiles;
    int64_t n_cells;
    uint8_t phase;                // 0 = left extension, 1 = right extension, 2 = done
    uint8_t first_tile;
    uint8_t prev_gap;             // was the previously scored column a gap column
    uint8_t anchor_gap;           // is the left part's column next to the anchor a gap column
    uint8_t left_any;             // left part has at least one column
    // current tile
    int32_t t_ref_len, t_query_len;
};

GactScheduler::GactScheduler(gact_engine *engine, const gact_params &params,
                             const std::vector<SeqView> &refs, const std::vector<SeqView> &reads,
                             const std::vector<SeqView> &reads_rc, int host_threads)
    : eng_(engine), p_(params), refs_(refs), reads_(reads), reads_rc_(reads_rc),
      threads_(std::max(1, host_threads)), pitch_(gact_engine_states_pitch_words(engine)) {}

static void fail(gact_engine *e, const char *what, int rc)
{
    throw std::runtime_error(std::string(what) + ": " + gact_status_string(rc) + ": " + gact_last_error(e));
}

// OpenMP keeps a persistent worker pool: the loop runs a few hundred times per shard with a
// body of ~100 us, so spawning std::threads per round would dominate
template <class F>
static void parallel_for(int threads, size_t n, F f)
{
    if (threads <= 1 || n < 512) { f(0, n); return; }
    const long chunks = (long)std::min<size_t>((size_t)threads * 4, (n + 127) / 128);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (long c = 0; c < chunks; c++) {
        const size_t a = n * (size_t)c / (size_t)chunks, b = n * (size_t)(c + 1) / (size_t)chunks;
        f(a, b);
    }
}

void GactScheduler::run(const std::vector<GactCall> &calls, std::vector<GactAlignment> &out, SchedulerStats *stats)
{
    const auto t_begin = std::chrono::steady_clock::now();
    const int T = p_.tile_size, thr = p_.first_tile_score_threshold;
    const int ma = p_.match, mi = p_.mismatch, go = p_.gap_open, ge = p_.gap_extend;
    const int max_batch = gact_engine_max_tiles(eng_);
    out.assign(calls.size(), GactAlignment{});

    std::vector<Active> act(calls.size());
    for (size_t k = 0; k < calls.size(); k++) {
        Active &a = act[k];
        memset(&a, 0, sizeof(a));
        a.call = (int32_t)k;
        a.ref_pos = a.rev_ref_pos = calls[k].ref_pos;
        a.query_pos = a.rev_query_pos = calls[k].query_pos;
        a.first_tile = 1;
        a.phase = 0;
    }

    auto ref_of = [&](const GactCall &c) -> const SeqView & { return refs_[c.ref_id]; };
    auto qry_of = [&](const GactCall &c) -> const SeqView & { return c.complement ? reads_rc_[c.query_id] : reads_[c.query_id]; };

    // Decide the next tile of a candidate (loop heads of gact.cpp:82 and :144); returns false when done.
    // `adv_ok`: the previous tile advanced in both sequences (or there was none yet in this phase).
    auto next_tile = [&](Active &a, bool adv_ok, gact_tile_desc &d) -> bool {
        const GactCall &c = calls[a.call];
        const int64_t RL = ref_of(c).len, QL = qry_of(c).len;
        if (a.phase >= 2) return false;
        if (a.phase == 0) {
            if (a.ref_pos > 0 && a.query_pos > 0 && (adv_ok || a.first_tile)) {
                a.t_ref_len = a.ref_pos > T ? T : a.ref_pos;
                a.t_query_len = a.query_pos > T ? T : a.query_pos;
                d.ref_off = gact_engine_seq_start(eng_, GACT_SET_REF, c.ref_id) + a.ref_pos - a.t_ref_len;
                d.query_off = gact_engine_seq_start(eng_, c.complement ? GACT_SET_READS_RC : GACT_SET_READS, c.query_id) +
                              a.query_pos - a.t_query_len;
                d.reverse = 0;
            } else {
                // gact.cpp:136-141: left part finished, restart from the first tile's maximum
                a.ab = a.ref_pos; a.bb = a.query_pos;
                a.ref_pos = a.rev_ref_pos; a.query_pos = a.rev_query_pos;
                a.phase = 1;
                // the right part's first column follows the left part's anchor-side column
                a.prev_gap = a.left_any ? a.anchor_gap : 0;
                adv_ok = true;
            }
        }
        if (a.phase == 1) {
            if (a.ref_pos < RL && a.query_pos < QL && (adv_ok || a.first_tile)) {
                a.t_ref_len = (a.ref_pos + T < RL) ? T : (int32_t)(RL - a.ref_pos);
                a.t_query_len = (a.query_pos + T < QL) ? T : (int32_t)(QL - a.query_pos);
                d.ref_off = gact_engine_seq_start(eng_, GACT_SET_REF, c.ref_id) + a.ref_pos;
                d.query_off = gact_engine_seq_start(eng_, c.complement ? GACT_SET_READS_RC : GACT_SET_READS, c.query_id) +
                              a.query_pos;
                d.reverse = 1;
            } else {
                a.phase = 2;
                return false;
            }
        }
        d.ref_len = a.t_ref_len;
        d.query_len = a.t_query_len;
        d.ref_set = GACT_SET_REF;
        d.query_set = c.complement ? GACT_SET_READS_RC : GACT_SET_READS;
        d.first = a.first_tile;
        d.reserved = 0;
        return true;
    };

    // Consume one tile result (bodies of the loops at gact.cpp:95-133 and :158-194 plus the
    // column scoring of :197-210).  Returns whether the tile advanced in both sequences.
    auto consume = [&](Active &a, const gact_tile_result &r, const uint32_t *st) -> bool {
        const GactCall &c = calls[a.call];
        const char *ref = ref_of(c).data, *qry = qry_of(c).data;
        const bool left = (a.phase == 0);
        a.n_tiles++;
        a.n_cells += (int64_t)a.t_ref_len * a.t_query_len;
        if (a.first_tile) {
            if (left) {
                a.ref_pos = a.ref_pos - a.t_ref_len + r.max_i;
                a.query_pos = a.query_pos - a.t_query_len + r.max_j;
                a.rev_ref_pos = a.ref_pos;
                a.rev_query_pos = a.query_pos;
            } else {
                a.ref_pos = a.ref_pos + a.t_ref_len - r.max_i;
                a.query_pos = a.query_pos + a.t_query_len - r.max_j;
            }
            a.first_tile_score = r.score;
            if (r.score < thr) {
                // `break` at gact.cpp:107-109 / :168-170: this phase ends, first_tile stays set
                if (left) {
                    a.ab = a.ref_pos; a.bb = a.query_pos;
                    a.ref_pos = a.rev_ref_pos; a.query_pos = a.rev_query_pos;
                    a.phase = 1;
                    a.prev_gap = 0;
                    return true;            // the right loop starts with i = j = tile_size (gact.cpp:142-143)
                }
                a.phase = 2;
                return false;
            }
        }
        const int n = r.n_states;
        if (n > 0) a.first_tile = 0;
        int score = a.score;
        int prev_gap = a.prev_gap;
        int ri = 0, qi = 0;                 // reference / query bases consumed so far in this tile
        for (int k = 0; k < n; k++) {
            const int s = (st[k >> 4] >> (2 * (k & 15))) & 3;
            bool gap;
            if (s == GACT_STATE_M) {
                const char rc = left ? ref[a.ref_pos - ri - 1] : ref[a.ref_pos + ri];
                const char qc = left ? qry[a.query_pos - qi - 1] : qry[a.query_pos + qi];
                gap = (rc == '-' || qc == '-');
                if (!gap) score += (qc == rc) ? ma : mi;
                ri++; qi++;
            } else {
                gap = true;
                if (s == GACT_STATE_I) ri++; else qi++;
            }
            if (gap) score += prev_gap ? ge : go;
            if (left && !a.left_any) { a.left_any = 1; a.anchor_gap = gap; }
            prev_gap = gap;
        }
        a.score = score;
        a.prev_gap = (uint8_t)prev_gap;
        if (left) { a.ref_pos -= r.i_steps; a.query_pos -= r.j_steps; }
        else      { a.ref_pos += r.i_steps; a.query_pos += r.j_steps; }
        // a first tile that passes the threshold but yields no state would spin forever in the
        // reference (only possible with a threshold <= 0); stop extending in this direction
        if (a.first_tile && n == 0) { a.first_tile = 0; return false; }
        return r.i_steps > 0 && r.j_steps > 0;
    };

    // Two groups of candidates ping-pong through two of the engine's in-flight slots.
    struct Group {
        std::vector<int32_t> ids;              // active candidates of this group
        std::vector<uint8_t> adv;              // advance flag from the previous tile
        std::vector<gact_tile_desc> descs;
        std::vector<int32_t> owner;            // descs[k] belongs to candidate owner[k]
        std::vector<gact_tile_result> res;
        std::vector<uint32_t> st;
        bool inflight = false;
    } grp[2];
    for (size_t k = 0; k < calls.size(); k++) grp[k & 1].ids.push_back((int32_t)k);
    for (auto &g : grp) g.adv.assign(g.ids.size(), 1);

    uint64_t tiles = 0, rounds = 0;
    auto build_and_submit = [&](Group &g) -> bool {
        // next tile of every candidate that is still alive (serial: cheap, and keeps ids ordered)
        std::vector<int32_t> alive;
        g.descs.clear(); g.owner.clear();
        alive.reserve(g.ids.size());
        for (size_t k = 0; k < g.ids.size(); k++) {
            gact_tile_desc d;
            memset(&d, 0, sizeof(d));
            Active &a = act[g.ids[k]];
            if (next_tile(a, g.adv[k] != 0, d)) {
                alive.push_back(g.ids[k]);
                g.descs.push_back(d);
                g.owner.push_back(g.ids[k]);
            }
        }
        g.ids.swap(alive);
        if (g.descs.empty()) return false;
        if ((int)g.descs.size() > max_batch) throw std::runtime_error("GactScheduler: more active candidates than max_tiles_per_batch");
        int rc = gact_engine_submit(eng_, (int)g.descs.size(), g.descs.data());
        if (rc) fail(eng_, "gact_engine_submit", rc);
        g.inflight = true;
        tiles += g.descs.size();
        rounds++;
        return true;
    };
    auto wait_and_consume = [&](Group &g) {
        const size_t n = g.descs.size();
        // zero-copy: read results and states straight from the engine's pinned buffers
        const gact_tile_result *res = nullptr;
        const uint32_t *st = nullptr;
        int got = 0;
        int rc = gact_engine_wait_view(eng_, &got, &res, &st);
        if (rc) fail(eng_, "gact_engine_wait_view", rc);
        if ((size_t)got != n) throw std::runtime_error("GactScheduler: batch size mismatch");
        g.inflight = false;
        g.adv.assign(n, 0);
        parallel_for(threads_, n, [&](size_t a, size_t b) {
            for (size_t k = a; k < b; k++)
                g.adv[k] = consume(act[g.owner[k]], res[k], st + k * (size_t)pitch_) ? 1 : 0;
        });
    };

    // split oversized groups: the engine bounds the batch size
    if ((int)std::max(grp[0].ids.size(), grp[1].ids.size()) > max_batch) {
        // process in waves of 2*max_batch candidates
        std::vector<GactCall> part;
        std::vector<GactAlignment> part_out;
        out.clear();
        for (size_t lo = 0; lo < calls.size(); lo += 2 * (size_t)max_batch) {
            const size_t hi = std::min(calls.size(), lo + 2 * (size_t)max_batch);
            part.assign(calls.begin() + lo, calls.begin() + hi);
            SchedulerStats ps;
            run(part, part_out, &ps);
            out.insert(out.end(), part_out.begin(), part_out.end());
            if (stats) { stats->tiles += ps.tiles; stats->cells += ps.cells; stats->rounds += ps.rounds; }
        }
        if (stats) stats->wall_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
        return;
    }

    // every round consumes at least one base of every live candidate, so the number of rounds is
    // bounded by the longest sequence; anything beyond that is a logic error, not a long alignment
    int64_t longest = 1;
    for (const auto &v : refs_) longest = std::max(longest, v.len);
    for (const auto &v : reads_) longest = std::max(longest, v.len);
    const uint64_t round_cap = 4 * (uint64_t)longest + 64;
    bool live0 = build_and_submit(grp[0]);
    bool live1 = build_and_submit(grp[1]);
    while (live0 || live1) {
        if (rounds > round_cap) throw std::runtime_error("GactScheduler: round limit exceeded (scheduler logic error)");
        if (live0) { wait_and_consume(grp[0]); }
        // group 1 (if any) is on the device while group 0 is consumed and resubmitted
        if (live0) live0 = build_and_submit(grp[0]);
        if (live1) { wait_and_consume(grp[1]); live1 = build_and_submit(grp[1]); }
    }

    uint64_t cells = 0;
    for (size_t k = 0; k < calls.size(); k++) {
        const Active &a = act[k];
        GactAlignment &o = out[k];
        o.ab = a.ab; o.bb = a.bb; o.ae = a.ref_pos; o.be = a.query_pos;
        o.score = a.score; o.first_tile_score = a.first_tile_score;
        o.n_tiles = a.n_tiles; o.n_cells = a.n_cells;
        cells += (uint64_t)a.n_cells;
    }
    if (stats) {
        stats->tiles += tiles; stats->cells += cells; stats->rounds += rounds;
        stats->wall_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    }
}

std::string format_overlap(const std::string &ref_name, const std::string &query_name,
                           const GactAlignment &a, bool complement)
{
    std::string s;
    s.reserve(128);
    s += "ref_id: "; s += ref_name;
    s += ", query_id: "; s += query_name;
    s += ", ab: "; s += std::to_string(a.ab);
    s += ", ae: "; s += std::to_string(a.ae);
    s += ", bb: "; s += std::to_string(a.bb);
    s += ", be: "; s += std::to_string(a.be);
    s += ", score: "; s += std::to_string(a.score);
    s += ", comp: "; s += complement ? "1" : "0";
    s += "\n";
    return s;
}

}  // namespace darwin
