// int_peak.cu -- integer / DPX issue-rate microbenchmark: the roofline
// denominator of the DP kernels (MEASURED_PEAKS.json only has HBM and bf16).
// Each thread runs 8 independent dependency chains of one instruction kind;
// the result is lane-operations per second over the whole chip.
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../include/gact_b200.h"

namespace {

template <int KIND>
__device__ __forceinline__ void step(int (&v)[8], int a, int b)
{
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (KIND == 0) asm volatile("{.reg .s32 t; add.s32 t, %0, %1; add.s32 %0, t, %2;}" : "+r"(v[k]) : "r"(v[(k + 1) & 7]), "r"(a));   // IADD3
        if (KIND == 1) v[k] = __vimax3_s32(v[k], a, b);
        if (KIND == 2) v[k] = __viaddmax_s32(v[k], a, b);
        if (KIND == 3) v[k] = __vimax3_s16x2(v[k], a, b);
        if (KIND == 4) v[k] = __viaddmax_s16x2(v[k], a, b);
        if (KIND == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[k]) : "r"(a), "r"(b));
        if (KIND == 6) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v[k]) : "r"(a), "r"(b));
        if (KIND == 7) { if (k & 1) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v[k]) : "r"(a), "r"(b)); else v[k] = __viaddmax_s32(v[k], a, b); }
        if (KIND == 8) { __half2 x = *reinterpret_cast<__half2 *>(&v[k]), y = *reinterpret_cast<__half2 *>(&a); v[k] = __hne2_mask(x, y); }   // HSET2
        if (KIND == 9) v[k] = __vadd2(v[k], a);                                                                   // VIADD.16x2
        if (KIND == 10) v[k] = __byte_perm(v[k], a, b);                                                           // PRMT
        if (KIND == 11) { if (k & 1) { __half2 x = *reinterpret_cast<__half2 *>(&v[k]), y = *reinterpret_cast<__half2 *>(&a); v[k] = __hne2_mask(x, y); } else v[k] = __vimax3_s16x2(v[k], a, b); }
        if (KIND == 12) v[k] = __shfl_up_sync(0xffffffffu, v[k], 1);                                             // SHFL
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) peak_kernel(int *out, int iters)
{
    int v[8];
    // operands come from memory so that they live in registers, not the constant bank
    const int a = out[1], b = out[2];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = threadIdx.x + k;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) step<KIND>(v, a, b);
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= v[k];
    if (s == 0x7fffffff) out[0] = s;
}

template <int KIND>
int run(int device, double *gops)
{
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GACT_ERR_CUDA;
    int *d = nullptr;
    if (cudaMalloc(&d, 16) != cudaSuccess) return GACT_ERR_NOMEM;
    { const int init[4] = {0, 3, -7, 0}; cudaMemcpy(d, init, 16, cudaMemcpyHostToDevice); }
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        peak_kernel<KIND><<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return GACT_ERR_CUDA; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    const double instr_per_thread = (double)iters * 16 * 8;
    *gops = instr_per_thread * blocks * threads / (best * 1e-3) / 1e9;
    return GACT_OK;
}

// latency: one warp, one dependent chain of the instruction; ns per instruction
template <int KIND>
__global__ void __launch_bounds__(32) lat_kernel(int *out, int iters)
{
    const int a = out[1], b = out[2];
    int v = threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 64; u++) {
            if (KIND == 2) v = __viaddmax_s32(v, a, b);
            if (KIND == 3) v = __vimax3_s16x2(v, a, b);
            if (KIND == 4) v = __viaddmax_s16x2(v, a, b);
            if (KIND == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(a), "r"(b));
            if (KIND == 6) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b));
            if (KIND == 10) v = __byte_perm(v, a, b);
            if (KIND == 12) v = __shfl_up_sync(0xffffffffu, v, 1);
            if (KIND == 13) { v = __viaddmax_s16x2(v, a, b); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(a), "r"(b)); }   // the D chain of the DP
        }
    }
    if (v == 0x7fffffff) out[0] = v;
}

template <int KIND>
int run_lat(int device, double *ns)
{
    if (cudaSetDevice(device) != cudaSuccess) return GACT_ERR_CUDA;
    int *d = nullptr;
    if (cudaMalloc(&d, 16) != cudaSuccess) return GACT_ERR_NOMEM;
    { const int init[4] = {0, 3, -7, 0}; cudaMemcpy(d, init, 16, cudaMemcpyHostToDevice); }
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        lat_kernel<KIND><<<1, 32>>>(d, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return GACT_ERR_CUDA; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *ns = (double)best * 1e6 / ((double)iters * 64 * (KIND == 13 ? 2 : 1));
    return GACT_OK;
}

}  // namespace

extern "C" int gact_int_peak(int device, int kind, double *gops_out)
{
    if (!gops_out) return GACT_ERR_ARG;
    switch (kind) {
        case 0: return run<0>(device, gops_out);
        case 1: return run<1>(device, gops_out);
        case 2: return run<2>(device, gops_out);
        case 3: return run<3>(device, gops_out);
        case 4: return run<4>(device, gops_out);
        case 5: return run<5>(device, gops_out);
        case 6: return run<6>(device, gops_out);
        case 7: return run<7>(device, gops_out);
        case 8: return run<8>(device, gops_out);
        case 9: return run<9>(device, gops_out);
        case 10: return run<10>(device, gops_out);
        case 11: return run<11>(device, gops_out);
        case 12: return run<12>(device, gops_out);
        // 100 + kind: latency of a dependent chain of that instruction, nanoseconds per instruction
        case 102: return run_lat<2>(device, gops_out);
        case 103: return run_lat<3>(device, gops_out);
        case 104: return run_lat<4>(device, gops_out);
        case 105: return run_lat<5>(device, gops_out);
        case 106: return run_lat<6>(device, gops_out);
        case 110: return run_lat<10>(device, gops_out);
        case 112: return run_lat<12>(device, gops_out);
        case 113: return run_lat<13>(device, gops_out);
        default: return GACT_ERR_ARG;
    }
}
