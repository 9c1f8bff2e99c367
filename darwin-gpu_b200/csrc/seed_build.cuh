// seed_build.cuh -- seed-position table construction on the GPU ("next" row 3).
//
// Device restatement of the reference's SeedPosTable constructor:
//   SeqToTwoBit()            ntcoding.cpp:87-103     (A/a C/c G/g T/t -> 0..3, anything else 0)
//   TwoBitToMinimizers()     ntcoding.cpp:126-153    (window minimum of Wang-hashed k-mers; an entry is
//                                                     appended when the minimum changes or w positions
//                                                     after the last entry; the LOOP position is stored)
//   SeedPosTable()           seed_pos_table.cpp:46-98 (sort by (hash, position); index_table_[s] = number
//                                                     of entries with hash <= s; pos_table_ = positions)
// The reference's scan carries (last minimizer, last position) from one position to the next.  Here the
// same entries come from a run formulation: change_p = (min_p != min_{p-1}), run start s_p = last
// position <= p with a change (0 before the first one), entry at p iff change_p or (p - s_p) % w == 0.
// Run starts cross thread blocks through a per-block "last change" value and a prefix maximum.
// The sort is the library radix sort (cub), the counterpart of the reference's __gnu_parallel::sort;
// the index is a histogram over the hashes followed by an inclusive prefix sum.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <string>

#include "dsoft.cuh"

namespace gact {

constexpr int SB_THREADS = 256;
constexpr int SB_PER_THREAD = 8;
constexpr int SB_CHUNK = SB_THREADS * SB_PER_THREAD;       // positions per block

struct SeedTableDev {
    uint32_t *d_index = nullptr;       // 4^k + 1 entries
    uint32_t *d_pos = nullptr;         // n_pos entries
    uint64_t index_entries = 0;
    uint32_t n_pos = 0;
    uint32_t max_occ = 0;              // kmer_max_occurence_ (seed_pos_table.cpp:57)
    int k = 0, w = 0;
    uint32_t bin_size = 0, ref_len = 0;
    double build_ms = -1.0;
};

__global__ void seed_pack_kernel(const uint8_t *__restrict__ raw, uint32_t len, uint32_t *__restrict__ packed, uint32_t n_words)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n_words; wi += stride) {
        uint32_t v = 0;
        const uint64_t b0 = (uint64_t)wi * 16;
#pragma unroll
        for (int x = 0; x < 16; x++) {
            const uint64_t idx = b0 + x;
            if (idx < len) {
                const int c = raw[idx] | 0x20;
                const uint32_t code = c == 'c' ? 1u : c == 'g' ? 2u : c == 't' ? 3u : 0u;
                v |= code << (2 * x);
            }
        }
        packed[wi] = v;
    }
}

__device__ __forceinline__ uint32_t sb_kmer(const uint32_t *__restrict__ two_bit, uint32_t pos, uint32_t mask)
{
    const uint32_t idx = pos >> 4, sh = 2 * (pos & 15);
    const uint64_t cat = ((uint64_t)__ldg(two_bit + idx + 1) << 32) | __ldg(two_bit + idx);
    return (uint32_t)(cat >> sh) & mask;
}

// Window minima and change flags of the block's positions [base, base + SB_CHUNK), base = w-1 + block * SB_CHUNK.
// sm_h: hashes of positions base-w .. base+SB_CHUNK-1 (index 0 = base-w).  Returns for this thread's
// SB_PER_THREAD consecutive positions the minima and a bit mask of changes.
__device__ __forceinline__ uint32_t sb_minima(const uint32_t *__restrict__ two_bit, uint32_t end, int k, int w, uint32_t base,
                                              uint32_t *sm_h, uint32_t (&m)[SB_PER_THREAD])
{
    const uint32_t mask = (1u << (2 * k)) - 1u;
    for (int x = threadIdx.x; x < SB_CHUNK + w; x += SB_THREADS) {
        const long long p = (long long)base - w + x;
        sm_h[x] = (p >= 0 && p < (long long)end) ? dsoft_hash32(sb_kmer(two_bit, (uint32_t)p, mask), mask) : 0xffffffffu;
    }
    __syncthreads();
    const int t0 = threadIdx.x * SB_PER_THREAD;
    // minimum of the window ending at position base + t0 - 1 (the predecessor of this thread's first position)
    uint32_t prev = 0xffffffffu;
    for (int x = 0; x < w; x++) prev = min(prev, sm_h[t0 + x]);          // positions base+t0-w .. base+t0-1
    if (base + t0 == (uint32_t)(w - 1)) prev = 0;                         // last_m starts at 0 (ntcoding.cpp:131)
    uint32_t changes = 0;
#pragma unroll
    for (int y = 0; y < SB_PER_THREAD; y++) {
        uint32_t v = 0xffffffffu;
        for (int x = 0; x < w; x++) v = min(v, sm_h[t0 + y + 1 + x]);      // positions p-w+1 .. p, p = base+t0+y
        m[y] = v;
        const bool in = (base + t0 + y) < end;
        if (in && v != prev) changes |= 1u << y;
        prev = v;
    }
    return changes;
}

// pass 1: last position with a change inside each block (or -1)
__global__ void __launch_bounds__(SB_THREADS)
seed_last_change_kernel(const uint32_t *__restrict__ two_bit, uint32_t end, int k, int w, long long *__restrict__ blk_last)
{
    __shared__ uint32_t sm_h[SB_CHUNK + 32];
    __shared__ long long sm_last;
    if (threadIdx.x == 0) sm_last = -1;
    const uint32_t base = (uint32_t)(w - 1) + blockIdx.x * (uint32_t)SB_CHUNK;
    uint32_t m[SB_PER_THREAD];
    const uint32_t changes = sb_minima(two_bit, end, k, w, base, sm_h, m);
    if (changes) {
        const long long last = (long long)base + threadIdx.x * SB_PER_THREAD + (31 - __clz(changes));
        atomicMax(&sm_last, last);
    }
    __syncthreads();
    if (threadIdx.x == 0) blk_last[blockIdx.x] = sm_last;
}

// exclusive prefix maximum over the blocks; the virtual run before the first change starts at 0
__global__ void __launch_bounds__(1024)
seed_carry_kernel(const long long *__restrict__ blk_last, long long *__restrict__ blk_carry, int n_blocks)
{
    __shared__ long long part[1024];
    const int per = (n_blocks + 1023) / 1024;
    const int a = min(n_blocks, (int)threadIdx.x * per), b = min(n_blocks, a + per);
    long long mx = -1;
    for (int x = a; x < b; x++) mx = max(mx, blk_last[x]);
    part[threadIdx.x] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int x = 0; x < 1024; x++) { const long long v = part[x]; part[x] = run; run = max(run, v); }
    }
    __syncthreads();
    long long run = part[threadIdx.x];
    for (int x = a; x < b; x++) { blk_carry[x] = run; run = max(run, blk_last[x]); }
}

// pass 2 (WRITE = false): entries per block.  pass 3 (WRITE = true): keys (hash << 32 | position) at the
// block's offset, histogram of the hashes into the index table.
template <bool WRITE>
__global__ void __launch_bounds__(SB_THREADS)
seed_emit_kernel(const uint32_t *__restrict__ two_bit, uint32_t end, int k, int w, const long long *__restrict__ blk_carry,
                 unsigned long long *__restrict__ blk_count, const unsigned long long *__restrict__ blk_offset,
                 unsigned long long *__restrict__ keys, uint32_t *__restrict__ hist)
{
    __shared__ uint32_t sm_h[SB_CHUNK + 32];
    __shared__ long long sm_warp[SB_THREADS / 32];
    __shared__ uint32_t sm_cnt[SB_THREADS / 32];
    const uint32_t base = (uint32_t)(w - 1) + blockIdx.x * (uint32_t)SB_CHUNK;
    uint32_t m[SB_PER_THREAD];
    const uint32_t changes = sb_minima(two_bit, end, k, w, base, sm_h, m);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long p0 = (long long)base + threadIdx.x * SB_PER_THREAD;
    // run start before this thread's first position: prefix maximum of "last change" over the earlier threads
    long long mine = changes ? p0 + (31 - __clz(changes)) : -1;
    long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = max(incl, t);
    }
    if (lane == 31) sm_warp[warp] = incl;
    __syncthreads();
    long long start = blk_carry[blockIdx.x];
    for (int x = 0; x < warp; x++) start = max(start, sm_warp[x]);
    const long long before = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane > 0) start = max(start, before);
    // entries of this thread
    uint32_t emit = 0;
    long long s = start;
#pragma unroll
    for (int y = 0; y < SB_PER_THREAD; y++) {
        const long long p = p0 + y;
        if (p < (long long)end) {
            if ((changes >> y) & 1u) { s = p; emit |= 1u << y; }
            else if ((p - s) % w == 0) emit |= 1u << y;
        }
    }
    const uint32_t cnt = __popc(emit);
    // block-level exclusive prefix of the counts
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sm_cnt[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, total = 0;
    for (int x = 0; x < SB_THREADS / 32; x++) { if (x < warp) woff += sm_cnt[x]; total += sm_cnt[x]; }
    if (!WRITE) {
        if (threadIdx.x == 0) blk_count[blockIdx.x] = total;
    } else {
        unsigned long long o = blk_offset[blockIdx.x] + woff + (inc - cnt);
#pragma unroll
        for (int y = 0; y < SB_PER_THREAD; y++)
            if ((emit >> y) & 1u) {
                keys[o++] = ((unsigned long long)m[y] << 32) | (unsigned long long)(p0 + y);
                atomicAdd(hist + m[y], 1u);
            }
    }
}

__global__ void seed_positions_kernel(const unsigned long long *__restrict__ keys, uint32_t n, uint32_t *__restrict__ pos)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) pos[i] = (uint32_t)keys[i];
}

inline void seed_table_free(SeedTableDev *t)
{
    if (t->d_index) cudaFree(t->d_index);
    if (t->d_pos) cudaFree(t->d_pos);
    *t = SeedTableDev();
}

// ref: HOST pointer to the concatenated, bin-padded reference (darwin.cpp:530-543).  Returns 0, or a
// non-zero code with *err set (1 bad argument, 2 out of memory, 3 CUDA error).
inline int seed_table_build_device(const char *ref, uint32_t ref_len, int k, uint32_t w, uint32_t occ_mult, uint32_t bin_size,
                                   cudaStream_t st, SeedTableDev *out, std::string *err)
{
    *out = SeedTableDev();
    if (!(k <= 15 && k > 3 && k > (int)w) || w == 0 || w > 32 || bin_size == 0 || !ref) {      // seed_pos_table.cpp:48-50
        *err = "seed_size/window_size out of range (3 < k <= 15, 0 < w < k, w <= 32)";
        return 1;
    }
    out->k = k; out->w = (int)w; out->bin_size = bin_size; out->ref_len = ref_len;
    out->max_occ = occ_mult * (1 + (ref_len >> (2 * k)));
    out->index_entries = ((uint64_t)1 << (2 * k)) + 1;
    const uint32_t n_words = 1 + ref_len / 16;                         // SeqToTwoBit allocation (ntcoding.cpp:88)
    const bool scan = 16ull * n_words >= (uint64_t)(k + (int)w);
    const uint32_t end = scan ? 16u * n_words - (uint32_t)k - w : 0u;  // loop bound of ntcoding.cpp:139
    const uint32_t n_scan = end > w - 1 ? end - (w - 1) : 0u;
    const int n_blocks = (int)((n_scan + SB_CHUNK - 1) / SB_CHUNK);

    uint8_t *d_raw = nullptr;
    uint32_t *d_two = nullptr;
    long long *d_last = nullptr, *d_carry = nullptr;
    unsigned long long *d_cnt = nullptr, *d_off = nullptr, *d_keys = nullptr, *d_keys2 = nullptr;
    void *d_temp = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = 0;
    auto bail = [&](int code, const std::string &msg) { rc = code; *err = msg; };
#define SB_CK(call, code)                                                                         \
    do {                                                                                          \
        cudaError_t _r = (call);                                                                  \
        if (_r != cudaSuccess && rc == 0) bail((code), std::string(#call) + ": " + cudaGetErrorString(_r)); \
    } while (0)
    do {
        SB_CK(cudaEventCreate(&e0), 3); SB_CK(cudaEventCreate(&e1), 3);
        SB_CK(cudaMalloc(&d_raw, std::max<size_t>(ref_len, 1)), 2);
        SB_CK(cudaMalloc(&d_two, ((size_t)n_words + 2) * 4), 2);
        SB_CK(cudaMalloc(&out->d_index, out->index_entries * 4), 2);
        SB_CK(cudaMalloc(&d_last, (size_t)std::max(n_blocks, 1) * 8), 2);
        SB_CK(cudaMalloc(&d_carry, (size_t)std::max(n_blocks, 1) * 8), 2);
        SB_CK(cudaMalloc(&d_cnt, ((size_t)n_blocks + 1) * 8), 2);
        SB_CK(cudaMalloc(&d_off, ((size_t)n_blocks + 1) * 8), 2);
        if (rc) break;
        SB_CK(cudaEventRecord(e0, st), 3);
        SB_CK(cudaMemcpyAsync(d_raw, ref, ref_len, cudaMemcpyHostToDevice, st), 3);
        SB_CK(cudaMemsetAsync(d_two, 0, ((size_t)n_words + 2) * 4, st), 3);
        SB_CK(cudaMemsetAsync(out->d_index, 0, out->index_entries * 4, st), 3);
        SB_CK(cudaMemsetAsync(d_cnt, 0, ((size_t)n_blocks + 1) * 8, st), 3);
        seed_pack_kernel<<<1184, 256, 0, st>>>(d_raw, ref_len, d_two, n_words);
        unsigned long long total = 0;
        size_t temp_bytes = 0;
        if (n_blocks > 0) {
            seed_last_change_kernel<<<n_blocks, SB_THREADS, 0, st>>>(d_two, end, k, (int)w, d_last);
            seed_carry_kernel<<<1, 1024, 0, st>>>(d_last, d_carry, n_blocks);
            seed_emit_kernel<false><<<n_blocks, SB_THREADS, 0, st>>>(d_two, end, k, (int)w, d_carry, d_cnt, nullptr, nullptr, nullptr);
            size_t tb = 0;
            SB_CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_cnt, d_off, n_blocks + 1, st), 3);
            SB_CK(cudaMalloc(&d_temp, std::max<size_t>(tb, 16)), 2);
            if (rc) break;
            SB_CK(cub::DeviceScan::ExclusiveSum(d_temp, tb, d_cnt, d_off, n_blocks + 1, st), 3);
            SB_CK(cudaMemcpyAsync(&total, d_off + n_blocks, 8, cudaMemcpyDeviceToHost, st), 3);
            SB_CK(cudaStreamSynchronize(st), 3);
            cudaFree(d_temp); d_temp = nullptr;
            if (rc) break;
        }
        if (total > 0xffffffffull) { bail(1, "more than 2^32 minimizers"); break; }
        out->n_pos = (uint32_t)total;
        SB_CK(cudaMalloc(&out->d_pos, std::max<size_t>(total, 1) * 4), 2);
        if (total > 0) {
            SB_CK(cudaMalloc(&d_keys, total * 8), 2);
            SB_CK(cudaMalloc(&d_keys2, total * 8), 2);
            if (rc) break;
            seed_emit_kernel<true><<<n_blocks, SB_THREADS, 0, st>>>(d_two, end, k, (int)w, d_carry, nullptr, d_off, d_keys, out->d_index);
            SB_CK(cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, d_keys, d_keys2, (unsigned long long)total, 0, 32 + 2 * k, st), 3);
            SB_CK(cudaMalloc(&d_temp, std::max<size_t>(temp_bytes, 16)), 2);
            if (rc) break;
            SB_CK(cub::DeviceRadixSort::SortKeys(d_temp, temp_bytes, d_keys, d_keys2, (unsigned long long)total, 0, 32 + 2 * k, st), 3);
            seed_positions_kernel<<<1184, 256, 0, st>>>(d_keys2, (uint32_t)total, out->d_pos);
            SB_CK(cudaStreamSynchronize(st), 3);
            cudaFree(d_temp); d_temp = nullptr;
            cudaFree(d_keys); d_keys = nullptr;
        }
        // index_table_[s] = number of entries with hash <= s: inclusive sum of the histogram, in place
        size_t tb = 0;
        SB_CK(cub::DeviceScan::InclusiveSum(nullptr, tb, out->d_index, out->d_index, (unsigned long long)out->index_entries, st), 3);
        SB_CK(cudaMalloc(&d_temp, std::max<size_t>(tb, 16)), 2);
        if (rc) break;
        SB_CK(cub::DeviceScan::InclusiveSum(d_temp, tb, out->d_index, out->d_index, (unsigned long long)out->index_entries, st), 3);
        SB_CK(cudaGetLastError(), 3);
        SB_CK(cudaEventRecord(e1, st), 3);
        SB_CK(cudaStreamSynchronize(st), 3);
        if (rc) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        out->build_ms = ms;
    } while (0);
#undef SB_CK
    cudaFree(d_raw); cudaFree(d_two); cudaFree(d_last); cudaFree(d_carry); cudaFree(d_cnt); cudaFree(d_off);
    cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_temp);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (rc) { cudaGetLastError(); seed_table_free(out); }
    return rc;
}

}  // namespace gact
