// legacy_adapter.cpp -- the reference's GPU boundary (gact.h:85-98: GPU_init / GPU_close /
// Align_Batch_GPU) implemented on top of include/gact_b200.h, so that the reference's own
// darwin.cpp + gact.cpp (-D GPU) link against libgact_b200.so unchanged.  This is the
// compatibility route of INTEGRATION.md (section B); it needs the reference's gact.h on the
// include path and is therefore only built where /root/reference exists (oracle/Makefile).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>
#include "gact.h"
#include "gact_b200.h"

static gact_params g_params;                       // GPU_init's arguments (cuda_host.cu:193)
struct CUDA_Stream_Holder { gact_engine *engine; };   // opaque in gact.h:49, ours to define

void GPU_init(int tile_size, int tile_overlap, int gap_open, int gap_extend, int match, int mismatch,
              int early_terminate, std::vector<GPU_storage> *s, int num_threads)
{
    g_params = gact_params{match, mismatch, gap_open, gap_extend, tile_size, tile_overlap, 0};
    for (int i = 0; i < num_threads; ++i) {        // one engine + stream per host thread, as cuda_host.cu:212-229
        s->push_back(GPU_storage());
        (*s)[i].stream = new CUDA_Stream_Holder{nullptr};
        if (gact_engine_create(&(*s)[i].stream->engine, 0, &g_params, BATCH_SIZE, nullptr) != GACT_OK) {
            fprintf(stderr, "%s\n", gact_last_error(nullptr)); exit(-1);   // the reference exits on CUDA errors
        }
    }
}

void GPU_close(std::vector<GPU_storage> *s, int num_threads)
{
    for (int i = 0; i < num_threads; ++i) { gact_engine_destroy((*s)[i].stream->engine); delete (*s)[i].stream; }
}

// The reference hands over the tile strings of every slot (numeric bases 0..3 after darwin.cpp:314-398);
// idle slots have ref_lens[t] == -1 (cuda_host.cu:70-73).  reverses[t] == 1 means "natural order" in the GPU
// build (cuda_host.cu:92); the library uses the CPU-build sense (align.cpp:130-131), hence the inversion.
int *Align_Batch_GPU(std::vector<std::string> ref_seqs, std::vector<std::string> query_seqs,
                     std::vector<int> ref_lens, std::vector<int> query_lens, int *sub_mat, int gap_open,
                     int gap_extend, std::vector<int> ref_poss, std::vector<int> query_poss,
                     std::vector<char> reverses, std::vector<char> firsts, int early_terminate, int tile_size,
                     GPU_storage *s, int num_blocks, int threads_per_block)
{
    gact_engine *e = s->stream->engine;
    const int B = num_blocks * threads_per_block, pitch = gact_engine_states_pitch_words(e);
    std::vector<const char *> rp, qp; std::vector<int64_t> rl, ql; std::vector<int> slot;
    for (int t = 0; t < B; ++t) if (ref_lens[t] != -1) {
        rp.push_back(ref_seqs[t].data()); rl.push_back(ref_lens[t]);
        qp.push_back(query_seqs[t].data()); ql.push_back(query_lens[t]); slot.push_back(t);
    }
    // bytes 0..3 -> 8-bit sets.  sub_mat is not consulted: the reference's own Align_Batch_GPU receives it and
    // never reads it either (cuda_host.cu:26 is its only mention; the kernel scores with the match/mismatch
    // constants of GPU_init), so base 4 ('N') against base 4 scores as a match on both sides.
    if (gact_engine_upload(e, GACT_SET_REF, rp.size(), rp.data(), rl.data()) != GACT_OK ||
        gact_engine_upload(e, GACT_SET_AUX, qp.size(), qp.data(), ql.data()) != GACT_OK) {
        fprintf(stderr, "%s\n", gact_last_error(e)); exit(-1);                 // the reference exits on CUDA errors
    }
    std::vector<gact_tile_desc> d(slot.size());
    for (size_t k = 0; k < slot.size(); ++k) {
        const int t = slot[k];
        d[k] = gact_tile_desc{gact_engine_seq_start(e, GACT_SET_REF, k), gact_engine_seq_start(e, GACT_SET_AUX, k),
                              ref_lens[t], query_lens[t], GACT_SET_REF, GACT_SET_AUX,
                              (uint8_t)(reverses[t] == 1 ? 0 : 1), (uint8_t)firsts[t], 0};
    }
    std::vector<gact_tile_result> r(slot.size()); std::vector<uint32_t> st(slot.size() * pitch);
    if (gact_engine_align_tiles(e, (int)d.size(), d.data(), r.data(), st.data()) != GACT_OK) {
        fprintf(stderr, "%s\n", gact_last_error(e)); exit(-1);
    }
    // back to the reference's result layout: stride 2*tile_size ints per slot,
    // [score, i_steps, j_steps, max_i, max_j, states..., -1]  (cuda_header.h:257-302, gact.cpp:434-473)
    int *out = (int *)malloc(sizeof(int) * B * 2 * tile_size);
    for (size_t k = 0; k < slot.size(); ++k) {
        int *o = out + 2 * tile_size * slot[k];
        o[0] = r[k].score; o[1] = r[k].i_steps; o[2] = r[k].j_steps; o[3] = r[k].max_i; o[4] = r[k].max_j;
        for (int x = 0; x < r[k].n_states; ++x) o[5 + x] = (st[k * pitch + (x >> 4)] >> (2 * (x & 15))) & 3;
        o[5 + r[k].n_states] = -1;
    }
    return out;
}
