// Fixed-order reductions of the per-CTA weight-gradient partials of the tensor-core conv kernels, shared by the
// single-layer entry points (conv_tc.cu, conv1_tc.cu) and the all-layers launch (multi.cu).
#pragma once
#include "common.cuh"

namespace drq {

constexpr int kWgPartialFloats = 10 * 32 * 32;     // conv3x3: [9 taps + bias][ci][co] per CTA
constexpr int kC1PartialFloats = 32 * 96;          // conv1: [co][96 K entries] per CTA

inline int conv_wgrad_ctas(int n_images, int hout) {             // tile walkers of conv3x3_wgrad_tc_kernel
    const int tiles = n_images * ((hout * DRQ_PW + 127) / 128);
    return tiles < sm_budget() ? tiles : sm_budget();
}
inline int conv1_wgrad_ctas(int n_images) {
    const int tiles = n_images * 14;
    return tiles < sm_budget() ? tiles : sm_budget();
}
constexpr int kWgReduceBlocks = (9248 + 31) / 32;
constexpr int kC1ReduceBlocks = 32 * 96 / 32;

// dw[co][ci][tap] = sum_g partial[g][tap][ci][co]; db[co] = sum_g partial[g][9][0][co].  Block = 32 outputs x 8
// slices of the G partials (fixed association), combined in fixed order.  256 threads.
__device__ __forceinline__ void wgrad3x3_reduce_block(const float* __restrict__ partial, int G, float* __restrict__ dw,
                                                      float* __restrict__ db, int bx, float (*red)[33]) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = bx * 32 + tx;
    const bool live = i < 9 * 32 * 32 + 32;
    const int src = i < 9216 ? i : 9 * 1024 + (i - 9216);
    float s = 0.f;
    if (live)
        for (int g = ty; g < G; g += 8) s += partial[(long long)g * kWgPartialFloats + src];
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && live) {
        float t = red[0][tx];
#pragma unroll
        for (int r = 1; r < 8; ++r) t += red[r][tx];
        if (i < 9216) {
            const int tap = i / 1024, ci = (i / 32) % 32, co = i % 32;
            dw[(co * 32 + ci) * 9 + tap] = t;
        } else {
            db[i - 9216] = t;
        }
    }
}

// conv1: partial[g][co][k] holds S = sum (x - 128) * d (k < cin*9) and sum d (k == cin*9);
// dW = S / 255 + (128/255 - 0.5) * db  (x/255 - 0.5 == (x - 128)/255 + (128/255 - 0.5)), db = sum d.
__device__ __forceinline__ void conv1_reduce_block(const float* __restrict__ partial, int G, int cin, float* __restrict__ dw,
                                                   float* __restrict__ db, int bx, float (*red)[33], float (*redb)[33]) {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = bx * 32 + tx;                         // 96 = 3 x 32: a block stays inside one co
    const int co = i / 96, k = i - co * 96;
    float s = 0.f, sb = 0.f;
    for (int g = ty; g < G; g += 8) {
        s += partial[(long long)g * kC1PartialFloats + i];
        sb += partial[(long long)g * kC1PartialFloats + co * 96 + cin * 9];
    }
    red[ty][tx] = s; redb[ty][tx] = sb;
    __syncthreads();
    if (ty == 0 && k <= cin * 9) {
        float t = red[0][tx], tb = redb[0][tx];
#pragma unroll
        for (int r = 1; r < 8; ++r) { t += red[r][tx]; tb += redb[r][tx]; }
        if (k < cin * 9) dw[co * cin * 9 + k] = fmaf(t, 1.0f / 255.0f, (128.0f / 255.0f - 0.5f) * tb); else db[co] = t;
    }
}

}  // namespace drq
