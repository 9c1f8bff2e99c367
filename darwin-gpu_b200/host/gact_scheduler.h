// gact_scheduler.h -- host-side GACT extension control for many candidates at once.
//
// Replaces the reference's GACT_Batch() slot scheduler (gact.cpp:231-560) and is
// observably equal to its CPU GACT() (gact.cpp:48-228): per candidate a left
// extension, a right extension from the first tile's maximum, the total score
// over the concatenated alignment columns, and the begin/end coordinates.  Every
// tile is one gact_tile_desc sent through the C ABI (include/gact_b200.h); the
// host never computes DP cells.
//
// Differences by design: all candidates of a shard are in flight at once (one
// tile per active candidate per round, thousands per launch) and are split in
// two groups so that the device aligns one group while the host consumes the
// other group's tracebacks; the aligned strings are never materialised -- the
// score recurrence of gact.cpp:197-210 is evaluated on the fly.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/gact_b200.h"

namespace darwin {

struct SeqView {
    const char *data;
    int64_t len;
};

struct GactCall {                // one D-SOFT candidate (darwin.cpp:215-246)
    int32_t ref_id, query_id;
    int32_t ref_pos, query_pos;
    uint8_t complement;          // 0: read as given, 1: reverse-complemented read
};

struct GactAlignment {           // what gact.cpp:213-225 prints
    int32_t ab, ae, bb, be;
    int32_t score;
    int32_t first_tile_score;
    int32_t n_tiles;
    int64_t n_cells;
};

struct SchedulerStats {
    uint64_t tiles = 0, cells = 0, rounds = 0, first_tiles = 0;
    double wall_ms = 0, device_ms = 0;
};

class GactScheduler {
public:
    // refs / reads / reads_rc: host views of the sequences uploaded to the engine's
    // GACT_SET_REF / GACT_SET_READS / GACT_SET_READS_RC sets, in the same order.
    GactScheduler(gact_engine *engine, const gact_params &params,
                  const std::vector<SeqView> &refs, const std::vector<SeqView> &reads,
                  const std::vector<SeqView> &reads_rc, int host_threads);

    // Extends every call; out[k] belongs to calls[k].  Throws std::runtime_error on engine errors.
    void run(const std::vector<GactCall> &calls, std::vector<GactAlignment> &out, SchedulerStats *stats = nullptr);

private:
    struct Active;
    gact_engine *eng_;
    gact_params p_;
    const std::vector<SeqView> &refs_, &reads_, &reads_rc_;
    int threads_;
    int pitch_;
};

// "ref_id: <r>, query_id: <q>, ab: .., ae: .., bb: .., be: .., score: .., comp: <0|1>\n"  (gact.cpp:214-224)
std::string format_overlap(const std::string &ref_name, const std::string &query_name,
                           const GactAlignment &a, bool complement);

}  // namespace darwin
