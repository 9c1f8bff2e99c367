/*
 * gact_b200.h -- C ABI of the B200-native GACT tile-alignment engine.
 *
 * This is the drop-in boundary that replaces the reference's CUDA host layer
 * (cuda_host.cu / cuda_header.h, declared in gact.h:48-98):
 *
 *   reference                                   this library
 *   ------------------------------------------  ---------------------------------
 *   GPU_init()          cuda_host.cu:193-237     gact_engine_create()
 *   GPU_close()         cuda_host.cu:239-258     gact_engine_destroy()
 *   (host strings copied per batch,              gact_engine_upload()  -- once,
 *    cuda_host.cu:85-163)                         2-bit packed, HBM-resident
 *   Align_Batch_GPU()   cuda_host.cu:23-190      gact_engine_align_tiles()
 *                                                gact_engine_submit()/_wait()
 *   GPU_storage         gact.h:51-67             opaque gact_engine
 *   int* result, stride 2*tile_size              gact_tile_result + 2-bit states
 *     cuda_header.h:257-302
 *   cudaSafeCall -> exit(-1)                     int status + gact_last_error()
 *     cuda_header.h:311-319
 *
 * Optional entry points beyond that boundary (callers and producers either side of the tile path):
 *   GACT() / GACT_Batch()  gact.cpp:48-228, 231-560        gact_engine_extend()
 *   SeedPosTable::DSOFT()  seed_pos_table.cpp:100-167      gact_dsoft_create()/_run()
 *   SeedPosTable()         seed_pos_table.cpp:46-98        gact_seed_table_build()
 *
 * Plain C types only; no CUDA, C++ or torch types cross this boundary.  All
 * functions return GACT_OK (0) or a negative gact_status; none of them calls
 * exit().  An engine is bound to one device and one stream and is meant to be
 * driven by one host thread (the reference's model: one GPU_storage + stream
 * per host thread, darwin.cpp:619-629); different engines are independent.
 *
 * Tile semantics are those of the reference CPU aligner AlignWithBT()
 * (align.cpp:60-233) bit for bit: clamped-M affine-gap recurrence, raw byte
 * equality, M>=I>=D tie order with the all-non-positive ZERO override, >= in
 * both gap flags, last maximum wins, early-terminate test before the push.
 * There is no CPU fallback anywhere behind this interface.
 */
#ifndef GACT_B200_H
#define GACT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GACT_B200_ABI_VERSION 1

typedef enum {
    GACT_OK = 0,
    GACT_ERR_ARG = -1,       /* bad argument / unsupported parameter combination */
    GACT_ERR_CUDA = -2,      /* a CUDA runtime call failed (see gact_last_error)  */
    GACT_ERR_NOMEM = -3,     /* host or device allocation failed                  */
    GACT_ERR_STATE = -4,     /* call sequence error (e.g. wait without submit)    */
    GACT_ERR_NODEVICE = -5   /* no usable CUDA device                             */
} gact_status;

/* params.cfg keys of the reference ([GACT_scoring], [GACT_extend],
 * [GACT_first_tile]; darwin.cpp:471-492).  gap_open and gap_extend must be
 * <= 0 (with a positive gap score the reference itself indexes dir[-1],
 * align.cpp:210-229).  tile_size <= GACT_MAX_TILE_SIZE, 0 <= tile_overlap <
 * tile_size.  first_tile_score_threshold is carried for the host scheduler;
 * the tile kernels do not use it. */
typedef struct {
    int32_t match;
    int32_t mismatch;
    int32_t gap_open;
    int32_t gap_extend;
    int32_t tile_size;
    int32_t tile_overlap;
    int32_t first_tile_score_threshold;
} gact_params;

#define GACT_MAX_TILE_SIZE 1024
#define GACT_MAX_SETS 4
#define GACT_MAX_INFLIGHT 3      /* async batches between gact_engine_submit() and gact_engine_wait() */

/* Sequence sets held on the device.  A set is one concatenated buffer; tiles
 * address it by base offset.  The three sets the reference path needs
 * (reference_seqs, reads_seqs, rev_reads_seqs; darwin.cpp:86-92) have names. */
enum { GACT_SET_REF = 0, GACT_SET_READS = 1, GACT_SET_READS_RC = 2, GACT_SET_AUX = 3 };

/* One tile = one AlignWithBT() call (align.h:29-32).
 *   ref_off/query_off : offset of the tile's FIRST base (lowest address) in
 *                       its set -- i.e. ref_str + ref_pos - ref_tile_length for
 *                       the left extension (gact.cpp:88) and ref_str + ref_pos
 *                       for the right extension (gact.cpp:150)
 *   reverse           : CPU-build sense of align.cpp:130-131 -- 0: bases are
 *                       consumed in natural order (left extension, gact.cpp:93),
 *                       1: back to front (right extension, gact.cpp:155)
 *   first             : traceback starts at the last maximum (align.cpp:190)
 */
typedef struct {
    int64_t ref_off;
    int64_t query_off;
    int32_t ref_len;      /* 0..tile_size */
    int32_t query_len;    /* 0..tile_size */
    uint8_t ref_set;
    uint8_t query_set;
    uint8_t reverse;
    uint8_t first;
    uint32_t reserved;    /* must be 0 */
} gact_tile_desc;         /* 32 bytes */

/* Per-tile result header.  Together with the packed states this carries
 * everything the reference's int result array does
 * ([score, i_steps, j_steps, max_i, max_j, states..., -1]; cuda_header.h:257-302). */
typedef struct {
    int32_t score;     /* first: max score; else H[ref_len][query_len]          */
    int32_t max_i;     /* traceback start, reference index (1-based); = ref_len
                          for non-first tiles                                    */
    int32_t max_j;     /* traceback start, query index (1-based)                */
    int32_t n_states;  /* number of traceback states (<= 2*(T-O) - 1)           */
    int32_t i_steps;   /* reference bases consumed (states M and I)             */
    int32_t j_steps;   /* query bases consumed     (states M and D)             */
} gact_tile_result;    /* 24 bytes */

/* States are 2 bits each (align.h:23: D=1, I=2, M=3), state k of a tile in
 * bits [2*(k%16), 2*(k%16)+1] of word k/16 of that tile's row; rows are
 * gact_engine_states_pitch_words() 32-bit words apart. */
#define GACT_STATE_D 1
#define GACT_STATE_I 2
#define GACT_STATE_M 3

typedef struct {
    uint64_t tiles;          /* tiles aligned since creation                     */
    uint64_t cells;          /* sum ref_len*query_len of those tiles             */
    uint64_t first_tiles;    /* tiles with first=1 (two-pass)                    */
    uint64_t kernel_launches;
    uint64_t batches;
    double   kernel_ms;      /* device time of the tile kernels (CUDA events)    */
    double   h2d_bytes;
    double   d2h_bytes;
} gact_stats;

typedef struct gact_engine gact_engine;

/* Library-level queries (no device needed). */
int         gact_abi_version(void);
const char *gact_status_string(int status);
int         gact_device_count(void);               /* <0: CUDA unavailable       */

/* Engine lifetime.  stream: a cudaStream_t passed as void*, or NULL to let the
 * engine create (and own) a non-blocking stream.  max_tiles_per_batch bounds n
 * in align_tiles/submit/stage. */
int  gact_engine_create(gact_engine **out, int device, const gact_params *params,
                        int max_tiles_per_batch, void *stream);
void gact_engine_destroy(gact_engine *e);
const char *gact_last_error(const gact_engine *e);  /* e may be NULL: last create error */

/* Upload one sequence set: n_seqs byte strings (raw FASTA characters, no
 * terminator needed), concatenated on the device in order.  Every set is kept
 * 2-bit packed (16 bases per 32-bit word; case folded, anything else 0 --
 * ntcoding.cpp:60-72, what D-SOFT hashes).  A set that holds any byte other
 * than 'A','C','G','T' (an "exception": N, lower case, IUPAC codes) also keeps
 * an exception bitmap and its raw bytes, so that the reference's raw byte
 * comparison (align.cpp:134: 'N'=='N', 'a'!='A') is preserved exactly: tiles
 * whose QUERY window holds an exception are aligned by the byte-comparing
 * kernels, all others (exceptions in the reference window included) by the
 * packed score-table kernels; the engine routes every tile of a batch itself.
 * Replaces a previous upload of the same set. */
int gact_engine_upload(gact_engine *e, int set, int64_t n_seqs,
                       const char *const *seqs, const int64_t *lens);
/* Offset of sequence i inside its set (what to add to a position to form
 * gact_tile_desc.ref_off / query_off). */
int64_t gact_engine_seq_start(const gact_engine *e, int set, int64_t i);
int64_t gact_engine_set_length(const gact_engine *e, int set);
int     gact_engine_set_bits(const gact_engine *e, int set);   /* 2: ACGT only, 8: raw bytes kept too, 0: empty */
/* 1 if sequence i of the set holds a byte other than ACGT, 0 if not, <0 on bad arguments.  Such a sequence cannot be
 * the QUERY of gact_engine_extend (its candidates go through the tile path); as a reference it can. */
int     gact_engine_seq_has_exceptions(const gact_engine *e, int set, int64_t i);

int gact_engine_states_pitch_words(const gact_engine *e);
int gact_engine_max_tiles(const gact_engine *e);

/* Synchronous batch: H2D descriptors, tile kernels, D2H results.  descs,
 * results and packed_states are HOST pointers (results: n entries;
 * packed_states: n * pitch words, may be NULL to skip the state download). */
int gact_engine_align_tiles(gact_engine *e, int n, const gact_tile_desc *descs,
                            gact_tile_result *results, uint32_t *packed_states);

/* Asynchronous pair over a ring of GACT_MAX_INFLIGHT internal slots: submit()
 * returns after queueing the copies and kernels of batch k; wait() blocks for
 * the OLDEST outstanding batch and copies its results out.  The batches run on
 * internal streams that start after the work already queued on the engine's
 * stream; the kernels of consecutive batches may overlap (the next batch fills
 * the SMs the previous one's last tiles leave idle). */
int gact_engine_submit(gact_engine *e, int n, const gact_tile_desc *descs);
int gact_engine_wait(gact_engine *e, gact_tile_result *results, uint32_t *packed_states);

/* Like gact_engine_wait(), but hands out pointers into the engine's pinned result buffers
 * instead of copying (n tiles; states rows gact_engine_states_pitch_words() apart).  The
 * pointers stay valid for (GACT_MAX_INFLIGHT - batches still in flight when this call returns)
 * further gact_engine_submit() calls. */
int gact_engine_wait_view(gact_engine *e, int *n, const gact_tile_result **results,
                          const uint32_t **packed_states);

/* Device-resident form (benchmarks, on-device pipelines): stage() copies the
 * descriptors once, run_staged() only launches the kernels (asynchronous on
 * the engine stream), fetch_staged() synchronises and copies results out. */
int gact_engine_stage(gact_engine *e, int n, const gact_tile_desc *descs);
int gact_engine_run_staged(gact_engine *e);
int gact_engine_fetch_staged(gact_engine *e, gact_tile_result *results, uint32_t *packed_states);
int gact_engine_sync(gact_engine *e);
/* Device time of the most recent run_staged()/align_tiles() kernels in ms
 * (CUDA events on the engine stream); <0 if none. */
double gact_engine_last_kernel_ms(gact_engine *e);

/* Diagnostics of the last finished tile batch: how many tiles the inter-task kernel took (full, non-first tiles of large
 * batches; one lane per pair of tiles, pointers only for a band around the diagonal) and how many of them it handed back
 * to the wavefront kernel because their traceback left that band.  Results never depend on the routing. */
int gact_engine_tile_path_info(const gact_engine *e, int *n_inter_task, int *n_handed_back);
/* Optional: allocate the batch slots of the tile path now (they are otherwise allocated by the first tile batch). */
int gact_engine_reserve_tiles(gact_engine *e);
int gact_engine_stats(const gact_engine *e, gact_stats *out);
int gact_engine_reset_stats(gact_engine *e);

/* Kernel selection (for tests and benchmarks): 0 = auto, 1 = int32 DPX kernel,
 * 2 = packed s16x2 DPX kernel.  Returns GACT_ERR_ARG if the variant cannot run
 * the engine's parameters. */
int gact_engine_set_kernel(gact_engine *e, int variant);
int gact_engine_get_kernel(const gact_engine *e);

/* ---- whole candidate extensions on the device (GACT(), gact.cpp:48-228) ----
 * One call = one D-SOFT candidate: left extension, right extension from the first tile's maximum,
 * first-tile threshold, total score -- the tile chain is walked on the GPU, the host gets one
 * gact_alignment per call and no traceback states.  Supported when the packed s16x2 score-table kernels can run the
 * engine's parameters (any tile_size <= GACT_MAX_TILE_SIZE with scores in the 16-bit range, |16 * score| < 128).
 * Bytes other than ACGT in the reference are handled; a call whose QUERY sequence holds one
 * (gact_engine_seq_has_exceptions) is refused with GACT_ERR_ARG -- the caller drives that candidate's tiles itself
 * (gact_engine_submit/wait, as host/gact_scheduler.cpp does). */
typedef struct {
    int32_t ref_seq;      /* sequence index inside GACT_SET_REF                    */
    int32_t query_seq;    /* sequence index inside query_set                       */
    int32_t ref_pos;      /* anchor (darwin.cpp:216-224)                           */
    int32_t query_pos;
    uint8_t query_set;    /* GACT_SET_READS or GACT_SET_READS_RC                   */
    uint8_t reserved[3];
} gact_call;

typedef struct {
    int32_t ab, ae, bb, be;       /* gact.cpp:219-222                               */
    int32_t score;                /* total score, gact.cpp:197-210                  */
    int32_t first_tile_score;
    int32_t n_tiles;
    int32_t reserved;
    int64_t n_cells;              /* sum ref_len*query_len over the tiles aligned   */
} gact_alignment;

int gact_engine_extend(gact_engine *e, int n, const gact_call *calls, gact_alignment *out);
/* Asynchronous form: up to GACT_MAX_INFLIGHT batches between submit and wait, each on its own stream, so the
 * chains of one batch of reads run while the next batch is being filtered (gact_dsoft_submit) and the previous
 * one is being written out.  wait returns the oldest outstanding batch, out[i] belonging to calls[i] of its submit. */
int gact_engine_extend_submit(gact_engine *e, int n, const gact_call *calls);
int gact_engine_extend_wait(gact_engine *e, gact_alignment *out);
/* Optional: allocate the device buffers for up to n calls per batch now (e.g. before a timed phase). */
int gact_engine_extend_reserve(gact_engine *e, int n);
int gact_engine_extend_supported(const gact_engine *e);      /* 1 / 0 */
/* How chains are mapped onto the GPU (for tests and benchmarks; the result never depends on it):
 * 0 = auto (from the number and expected length of the chains), 1 = latency kernel (one tile per warp, direction
 * window in shared memory, tile_size <= 320) with one warp per SM sub-partition, 2 = the same with two,
 * 3 = throughput kernel (two tiles per warp for tile_size <= 320), 4 = the longest chains on the latency kernel,
 * the rest on the throughput kernel.  gact_engine_chain_info reports what the last submit used. */
int gact_engine_set_chain_mode(gact_engine *e, int mode);
int gact_engine_chain_info(const gact_engine *e, int *mode, int *ctas, int *n_long);

/* ---- D-SOFT candidate filter on the device (seed_pos_table.cpp:100-167, ntcoding.cpp:155-182) ----
 * The seed-position table (index_table_: 4^k + 1 entries, pos_table_) is built by the host
 * (SeedPosTable constructor, seed_pos_table.cpp:46-98) and uploaded once; queries are sequences of
 * the engine's uploaded sets.  Candidates come back grouped by query, in the reference's emission
 * order: (hit << 32) | offset of the reference is {hit, offset} here. */
typedef struct gact_dsoft gact_dsoft;
typedef struct {
    int32_t  query;     /* index into the queries passed to gact_dsoft_run               */
    int32_t  seq;       /* emission order inside that query                              */
    uint32_t hit;       /* position in the concatenated, bin-padded reference            */
    uint32_t offset;    /* position in the query                                         */
} gact_dsoft_cand;

int  gact_dsoft_create(gact_dsoft **out, gact_engine *e, const uint32_t *index_table, uint64_t index_entries,
                       const uint32_t *pos_table, uint64_t n_pos, int kmer_size, int window_size,
                       uint32_t bin_size, uint32_t kmer_max_occurence, int num_seeds, int threshold,
                       int max_candidates);
void gact_dsoft_destroy(gact_dsoft *d);

/* Seed-position table built on the device: replaces the SeedPosTable constructor
 * (seed_pos_table.cpp:46-98: SeqToTwoBit ntcoding.cpp:87-103, TwoBitToMinimizers ntcoding.cpp:126-153,
 * sort, index fill).  ref is the HOST string darwin.cpp:530-543 builds: every reference sequence
 * padded with 'N' to a multiple of bin_size, concatenated.  The tables stay in device memory;
 * gact_dsoft_create_from_table() borrows them (destroy the filter before the table),
 * gact_seed_table_download() copies them out (index_table: 4^k + 1 words, pos_table: num_minimizers words;
 * either may be NULL). */
typedef struct gact_seed_table gact_seed_table;
int  gact_seed_table_build(gact_seed_table **out, gact_engine *e, const char *ref, uint32_t ref_len, int kmer_size,
                           uint32_t seed_occurence_multiple, uint32_t bin_size, uint32_t window_size);
void gact_seed_table_destroy(gact_seed_table *t);
int  gact_seed_table_info(const gact_seed_table *t, uint64_t *index_entries, uint32_t *num_minimizers,
                          uint32_t *kmer_max_occurence, double *build_ms);
int  gact_seed_table_download(const gact_seed_table *t, uint32_t *index_table, uint32_t *pos_table);
int  gact_dsoft_create_from_table(gact_dsoft **out, gact_engine *e, const gact_seed_table *t, int num_seeds,
                                  int threshold, int max_candidates);
/* queries: (set, sequence index inside the set) pairs.  Returns GACT_ERR_NOMEM with *n_out set to the
 * required capacity when out_cap is too small. */
int  gact_dsoft_run(gact_dsoft *d, int n_queries, const int32_t *sets, const int64_t *seq_index,
                    gact_dsoft_cand *out, int64_t out_cap, int64_t *n_out);
/* Asynchronous form on the filter's own stream (two batches in flight at most): submit enqueues the queries, the
 * kernel and the copy of up to out_cap candidates; wait returns the oldest batch's candidates like gact_dsoft_run. */
int  gact_dsoft_submit(gact_dsoft *d, int n_queries, const int32_t *sets, const int64_t *seq_index, int64_t out_cap);
int  gact_dsoft_wait(gact_dsoft *d, gact_dsoft_cand *out, int64_t out_cap, int64_t *n_out);
double gact_dsoft_last_kernel_ms(const gact_dsoft *d);
/* Optional: allocate the device buffers for n_queries queries / out_cap candidates now. */
int  gact_dsoft_reserve(gact_dsoft *d, int n_queries, int64_t out_cap);

/* Integer/DPX issue-rate microbenchmark (the roofline denominator of the DP
 * kernels; MEASURED_PEAKS.json has none).  kind: 0 = IADD3, 1 = VIMNMX3.S32,
 * 2 = VIADDMNMX.S32, 3 = VIMNMX3.S16x2, 4 = VIADDMNMX.S16x2, 5 = LOP3,
 * 6 = IMAD, 7 = 1:1 VIADDMNMX:IMAD mix, 8 = HSET2, 9 = VIADD.16x2, 10 = PRMT,
 * 11 = 1:1 VIMNMX3.S16x2:HSET2 mix, 12 = SHFL.  Writes giga lane-operations
 * per second (one per thread per instruction).  kind 100 + k (k = 2, 3, 4, 5, 6, 10, 12; 113 = the
 * VIADDMNMX.S16x2 + LOP3 pair of the DP's D chain): latency of a dependent chain of that instruction on one
 * warp, written as nanoseconds per instruction. */
int gact_int_peak(int device, int kind, double *gops_out);

#ifdef __cplusplus
}
#endif
#endif /* GACT_B200_H */
