// Device bodies of the fp32 -> bf16 operand re-pack kernels, shared by the single-tensor entry points
// (gemm_tc.cu, conv_tc.cu, conv1_tc.cu) and the multi-job launcher (multi.cu): after an optimiser step
// all bf16 copies of a network are refreshed by ONE launch instead of one per tensor.
#pragma once
#include "tc_common.cuh"

namespace drq {

// element offset of TB element (row, unit) with R rows per block
__device__ __forceinline__ long long tb_off(long long row, int unit, int units, int R) {
    return (((row / R) * units + unit) * R + (row % R)) * 8;
}

// fp32 nn.Linear weight [rows][cols] -> TB(64) bf16; block (bx, u) = 256 rows of K unit u
__device__ __forceinline__ void pack_linear_tb_block(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int rows,
                                                     int cols, int units, int bx, int u, int tid) {
    const int r = bx * 256 + tid;
    const int rpad = (rows + DRQ_TB_W - 1) / DRQ_TB_W * DRQ_TB_W;
    if (r >= rpad) return;
    uint32_t pk[4] = {0, 0, 0, 0};
    if (r < rows) {
        const float* src = w + (long long)r * cols + u * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = u * 8 + 2 * j;
            pk[j] = tc::pack_bf16x2(c < cols ? src[2 * j] : 0.f, c + 1 < cols ? src[2 * j + 1] : 0.f);
        }
    }
    *reinterpret_cast<uint4*>(out + tb_off(r, u, units, DRQ_TB_W)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// trunk weight fp32 [rows][32*1225] (reference NCHW-flatten columns c*1225+yx) -> TB(64) bf16 with the feature
// order of the bf16 encoder output, n' = (c/8)*9800 + yx*8 + c%8: unit (c/8)*1225 + yx.  Block (bx, r): 32 pixels of weight row r, 32x32 transpose in `tile`.
template <int R = DRQ_TB_W>
__device__ __forceinline__ void pack_trunk_tb_block(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int rows,
                                                    int bx, int r, int tid, float (*tile)[33]) {
    const int yx0 = bx * 32;
    const int tx = tid & 31, ty = tid >> 5;
    const float* wr = w + (long long)r * DRQ_REPR_DIM;
    const bool live = r < rows;
#pragma unroll
    for (int c = ty; c < 32; c += 8) {
        const int yx = yx0 + tx;
        tile[c][tx] = (live && yx < 1225) ? wr[c * 1225 + yx] : 0.f;
    }
    __syncthreads();
    if (tid < 128) {       // thread -> (yx = yx0 + i, channel unit cu): 32 x 4 = 128 units per tile
        const int i = tid >> 2, cu = tid & 3;
        const int yx = yx0 + i;
        if (yx < 1225) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) pk[j] = tc::pack_bf16x2(tile[cu * 8 + 2 * j][i], tile[cu * 8 + 2 * j + 1][i]);
            *reinterpret_cast<uint4*>(out + tb_off(r, cu * 1225 + yx, DRQ_REPR_DIM / 8, R)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}

// fp32 conv weights [co][ci][3][3] -> bf16 UMMA B operands [tap*4 + k/8][n][k%8]:
//   fwd: n = co, k = ci (out = in * W);  dgrad: n = ci, k = co (din = dout * W^T with flipped offsets)
__device__ __forceinline__ void pack_conv_w_elem(const float* __restrict__ w, __nv_bfloat16* __restrict__ w_fwd,
                                                 __nv_bfloat16* __restrict__ w_dgrad, int i) {
    if (i >= 32 * 32 * 9) return;
    const int co = i / 288, ci = (i / 9) % 32, tap = i % 9;
    const __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
    w_fwd[((tap * 4 + ci / 8) * 32 + co) * 8 + (ci & 7)] = v;
    w_dgrad[((tap * 4 + co / 8) * 32 + ci) * 8 + (co & 7)] = v;
}

// fp32 conv1 weight [32][cin][3][3] + bias -> bf16 [12 K units][32 co][8] (zero padded to K = 96) followed by the
// fused forward bias b'[co] = b[co] + (128/255 - 0.5) * sum_k bf16(W[co][k])   (one block of 32 x 8 threads)
__device__ __forceinline__ void pack_conv1_w_block(const float* __restrict__ w, const float* __restrict__ bias,
                                                   __nv_bfloat16* __restrict__ out, int cin, int tid, float (*part)[9]) {
    const int co = tid >> 3, e = tid & 7;
    float s = 0.f;
    for (int u = 0; u < 12; ++u) {
        const int k = u * 8 + e;
        const __nv_bfloat16 h = __float2bfloat16_rn(k < cin * 9 ? w[co * cin * 9 + k] : 0.f);
        out[(u * 32 + co) * 8 + e] = h;
        s += __bfloat162float(h);
    }
    part[co][e] = s;
    __syncthreads();
    if (e == 0) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += part[co][j];
        reinterpret_cast<float*>(out + 12 * 32 * 8)[co] = bias[co] + (128.0f / 255.0f - 0.5f) * t;
    }
}

}  // namespace drq
