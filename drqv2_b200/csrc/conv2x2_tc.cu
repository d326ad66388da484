// 3x3 stride-1 convolutions of Encoder (drqv2.py:56-59), forward and data gradient, as implicit GEMMs whose accumulator
// row is a 2x2 BLOCK of output pixels (tcgen05 + TMEM, sm_100a only).
//
// Why: a tcgen05.mma M128 K16 costs 44.6 cycles at N = 32, 48.1 at N = 64 and 64.1 at N = 128 (tools/ub/ub_mma.cu), so
// the one-pixel-per-row kernel of conv_tc.cu (N = 32 output channels, 18 UMMAs per 128 pixels = 6.3 cycles per pixel)
// cannot pass 36 % of the tensor peak.  Here a row is the block of output pixels (2i+oy, 2j+ox), its 128 accumulator
// columns are (oy, ox, co), and the contraction runs over the block's 4x4 input window: window element (wy, wx)
// multiplies the weights of tap (wy-oy, wx-ox) of every output of the block it reaches.  The four inner window elements
// reach all four outputs (N = 128, no padding), the top / bottom edges two adjacent column groups (N = 64), the left /
// right edges two non-adjacent groups (N = 128 with zero weights), the corners one (N = 32): 32 UMMAs per 128 blocks
// = 512 pixels, 1768 cycles = 3.45 cycles per pixel.
//
// The activations stay in the WB layout (conv_tc.cu): the loader warps split a tile's input window into its four
// parity planes (pixel (2i+py, 2j+px) -> plane (py, px), block position i*pitch + j) with 16-byte cp.async copies, so
// that window element (wy, wx) of 128 consecutive blocks is plane (wy&1, wx&1) at a constant row offset
// (wy>>1)*pitch + (wx>>1): one K-major no-swizzle descriptor plus an offset per UMMA, as in conv_tc.cu.  Copies whose
// source pixel lies outside the valid input are zero-filled (cp.async src-size 0), so the data gradient needs no guard
// rows.  Block positions are numbered across images (pitch*nrow per image), so tiles are full but for the last one.
//
// The weights arrive in the compact operand layout of pack_conv_w_elem ([tap*4 + k/8][n][k%8], 18 KB) and are expanded
// per CTA into the 16 per-window-element B operands (88 KB of shared memory) while the loaders fill the first stages.
//
// Warp roles (416 threads, one CTA per SM): warp 0 = UMMA issuer, warps 1..4 = loaders, warps 5..12 = epilogue (lane
// quarter = warp % 4, output row parity oy = (warp - 5) / 4).
#include "tc_common.cuh"

namespace drq {

using namespace tc;

namespace c2 {

constexpr int kPLB = DRQ_PLB;
constexpr int kGuard = DRQ_GUARD;
constexpr int kSlack = DRQ_WB_SLACK;
constexpr int kTile = 128;                 // block positions per tile
constexpr int kWS = 152;                   // window slots per (plane, channel block): 128 + pitch + 1 <= 150
constexpr int kRegion = kWS * 16;          // bytes of one (plane, channel block)
constexpr int kStageBytes = 16 * kRegion;  // 4 planes x 4 channel blocks
constexpr int kStages = 3;
constexpr int kAcc = 4;                    // accumulator ring: 4 x 128 TMEM columns
constexpr int kWUnits = 8 * 512 + 4 * 256 + 4 * 128;   // 16-byte units of the expanded weights
constexpr int kWBytes = kWUnits * 16;
constexpr int kThreads = 13 * 32;
constexpr int kLoaders = 128;
constexpr int kExpanders = kThreads - kLoaders;
constexpr int kMaxPitch = 21;

struct Args {
    const __nv_bfloat16* in; long long cs_in;       // chunk (channel block) stride in pixel rows
    const __nv_bfloat16* w;                         // compact [36][32][8]
    const float* bias;                              // fwd
    const __nv_bfloat16* mask; long long cs_mask;   // dgrad: the layer's input activation (WB)
    __nv_bfloat16* out; long long cs_out;
    int n_images, pitch, pl4, total_pos, total_tiles;
    int h_in;                                       // valid height = width of the input; outside reads as zero
    int h_out;                                      // valid height = width of the output
    int out_mode;                                   // 0 WB, 1 compact NHWC, 2 TB features
    uint32_t m_pl4, m_pitch;                        // floor(2^32 / d) + 1: x / d = umulhi(x, m) for x * d < 2^32
    long long feat_rpad; int feat_half, feat_half_row;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// window element of B region r (regions ordered by UMMA shape: 4 inner, 4 left/right, 4 top/bottom, 4 corners)
__host__ __device__ constexpr int region_wy(int r) {
    return r < 4 ? 1 + (r >> 1) : r < 8 ? 1 + ((r - 4) >> 1) : r < 12 ? 3 * ((r - 8) >> 1) : 3 * ((r - 12) >> 1);
}
__host__ __device__ constexpr int region_wx(int r) {
    return r < 4 ? 1 + (r & 1) : r < 8 ? 3 * ((r - 4) & 1) : r < 12 ? 1 + ((r - 8) & 1) : 3 * ((r - 12) & 1);
}
__host__ __device__ constexpr int region_n(int r) { return r < 8 ? 128 : r < 12 ? 64 : 32; }
__host__ __device__ constexpr int region_unit0(int r) { return r < 8 ? r * 512 : r < 12 ? 4096 + (r - 8) * 256 : 5120 + (r - 12) * 128; }
// first accumulator column the region's UMMA writes
__host__ __device__ constexpr int region_col(int r) {
    return r < 8 ? 0 : r < 12 ? (region_wy(r) / 3) * 64 : ((region_wy(r) / 3) * 2 + region_wx(r) / 3) * 32;
}

template <bool DGRAD>
__device__ __forceinline__ void expand_weights(const __nv_bfloat16* __restrict__ w, uint8_t* w_s, int e) {
    const uint4* src = reinterpret_cast<const uint4*>(w);
    for (int u = e; u < kWUnits; u += kExpanders) {
        int r, ku, n;
        if (u < 4096) { r = u >> 9; ku = (u >> 7) & 3; n = u & 127; }
        else if (u < 5120) { const int x = u - 4096; r = 8 + (x >> 8); ku = (x >> 6) & 3; n = x & 63; }
        else { const int x = u - 5120; r = 12 + (x >> 7); ku = (x >> 5) & 3; n = x & 31; }
        const int wy = region_wy(r), wx = region_wx(r);
        int oy, ox;
        if (r < 8) { oy = n >> 6; ox = (n >> 5) & 1; }
        else if (r < 12) { oy = wy / 3; ox = n >> 5; }
        else { oy = wy / 3; ox = wx / 3; }
        const int co = n & 31;
        const int dy = DGRAD ? oy + 2 - wy : wy - oy;
        const int dx = DGRAD ? ox + 2 - wx : wx - ox;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (dy >= 0 && dy <= 2 && dx >= 0 && dx <= 2) val = __ldg(src + ((dy * 3 + dx) * 4 + ku) * 32 + co);
        reinterpret_cast<uint4*>(w_s)[u] = val;
    }
}

template <bool DGRAD>
__global__ void __launch_bounds__(kThreads, 1) conv2x2_tc_kernel(const Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* w_s = smem;
    uint8_t* a_s = smem + kWBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_s + kStages * kStageBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kStages;
    uint64_t* tfull = bars + 2 * kStages;
    uint64_t* tempty = bars + 2 * kStages + kAcc;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAcc);
    float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 4); mbar_init(empty + i, 1); }
        for (int i = 0; i < kAcc; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 8); }
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, kAcc * 128);
        tmem_relinquish();
    }
    pdl_wait();
    if (!DGRAD && threadIdx.x < 32) bias_s[threadIdx.x] = a.bias[threadIdx.x];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int pitch = a.pitch;

    if (warp >= 1 && warp <= 4) {
        // ------------------------------------------------ loaders: WB pixel rows -> four parity planes per channel block
        const int lt = threadIdx.x - 32;
        int stage = 0; uint32_t phase = 0; int prev = -1;
        const int n_slots = kTile + pitch + 1;
        // one window slot: decode its block position, then copy the (row parity, channel block) pairs PC0..PC0+NPC-1
        auto decode = [&](int v, int& n, int& i, int& j) -> bool {
            const bool inb = v >= 0 && v < a.total_pos;
            n = 0; i = 0; j = 0;
            if (inb) {
                n = (int)__umulhi((uint32_t)v, a.m_pl4);
                const int q = v - n * a.pl4;
                i = (int)__umulhi((uint32_t)q, a.m_pitch);
                j = q - i * pitch;
            }
            return inb;
        };
        auto copy_pc = [&](uint32_t sbase, bool inb, int n, int i, int j, int s, int pc) {
            const int py = pc >> 2, c = pc & 3;
            const int y = 2 * i + py, x0 = 2 * j;
            const bool rowok = inb && y < a.h_in;
            const bool ok0 = rowok && x0 < a.h_in, ok1 = rowok && x0 + 1 < a.h_in;
            const __nv_bfloat16* src = a.in + (c * a.cs_in + (long long)n * kPLB + kGuard + y * kPW + x0) * 8;
            const uint32_t dst = sbase + (uint32_t)(((py * 8 + c) * kWS + s) * 16);
            cp_async16(dst, ok0 ? src : a.in, ok0 ? 16u : 0u);
            cp_async16(dst + 4 * kRegion, ok1 ? src + 8 : a.in, ok1 ? 16u : 0u);
        };
        for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
            mbar_wait(empty + stage, phase ^ 1);
            const int v0 = t * kTile - (DGRAD ? pitch + 1 : 0);
            const uint32_t sbase = smem_u32(a_s + stage * kStageBytes);
            {                                               // slots 0..127: one per thread, all (py, channel block)
                int n, i, j;
                const bool inb = decode(v0 + lt, n, i, j);
#pragma unroll
                for (int pc = 0; pc < 8; ++pc) copy_pc(sbase, inb, n, i, j, lt, pc);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {                   // slots 128..151: (pc, slot) pairs spread over the threads
                const int idx = lt + k * kLoaders;
                const int pc = idx / 24, s = kTile + idx - pc * 24;
                if (pc < 8 && s < n_slots) {
                    int n, i, j;
                    const bool inb = decode(v0 + s, n, i, j);
                    copy_pc(sbase, inb, n, i, j, s, pc);
                }
            }
            cp_async_commit();
            if (prev >= 0) {                                // the previous tile's copies have landed: hand it to the UMMA warp
                cp_async_wait<1>();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(full + prev);
            }
            prev = stage;
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (prev >= 0) {
            cp_async_wait<0>();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full + prev);
        }
        pdl_release();
    } else {
        // ------------------------------------------------ everyone else first expands the weights
        expand_weights<DGRAD>(a.w, w_s, warp == 0 ? lane : threadIdx.x - kLoaders);
        fence_proxy_async();
        asm volatile("bar.sync 1, %0;" ::"n"(kExpanders) : "memory");
        if (warp == 0) {
            // -------------------------------------------- UMMA issuer
            constexpr uint32_t idesc128 = make_idesc_bf16(128, 128, false, false);
            constexpr uint32_t idesc64 = make_idesc_bf16(128, 64, false, false);
            constexpr uint32_t idesc32 = make_idesc_bf16(128, 32, false, false);
            const uint64_t da0 = make_smem_desc(smem_u32(a_s), kRegion, 128);
            const uint64_t db128 = make_smem_desc(smem_u32(w_s), 128 * 16, 128);
            const uint64_t db64 = make_smem_desc(smem_u32(w_s), 64 * 16, 128);
            const uint64_t db32 = make_smem_desc(smem_u32(w_s), 32 * 16, 128);
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
                mbar_wait(tempty + acc, acc_phase ^ 1);
                mbar_wait(full + stage, phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da_row0 = da0 + (uint64_t)(stage * (kStageBytes >> 4));
                    const uint64_t da_row1 = da_row0 + (uint64_t)pitch;
                    const uint32_t d_tmem = tmem_base + acc * 128;
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const int wy = region_wy(r), wx = region_wx(r);
                        const int plane = (wy & 1) * 2 + (wx & 1);
                        const uint64_t da = ((wy >> 1) ? da_row1 : da_row0) + (uint64_t)(plane * 4 * kWS + (wx >> 1));
                        const int nt = region_n(r);
                        const uint64_t db = (nt == 128 ? db128 : nt == 64 ? db64 : db32) + (uint64_t)region_unit0(r);
                        const uint32_t idesc = nt == 128 ? idesc128 : nt == 64 ? idesc64 : idesc32;
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            umma_bf16(d_tmem + region_col(r), da + (uint64_t)(h * 2 * kWS), db + (uint64_t)(h * 2 * nt), idesc,
                                      (r | h) ? 1u : 0u);
                    }
                    umma_commit(empty + stage);
                    umma_commit(tfull + acc);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
                if (++acc == kAcc) { acc = 0; acc_phase ^= 1; }
            }
        } else {
            // -------------------------------------------- epilogue: TMEM lane quarter q, output rows of parity oy
            const int q = warp & 3, oy = (warp - 5) >> 2;
            int acc = 0; uint32_t acc_phase = 0;
            float bias_r[DGRAD ? 1 : 32];
            if (!DGRAD) {
#pragma unroll
                for (int i = 0; i < 32; ++i) bias_r[i] = bias_s[i];
            }
            for (int t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
                const int v = t * kTile + q * 32 + lane;
                int n = 0, i = 0, j = 0;
                if (v < a.total_pos) {
                    n = (int)__umulhi((uint32_t)v, a.m_pl4);
                    const int qq = v - n * a.pl4;
                    i = (int)__umulhi((uint32_t)qq, a.m_pitch);
                    j = qq - i * pitch;
                }
                const int y = 2 * i + oy;
                const bool rowok = v < a.total_pos && y < a.h_out;
                const long long row0 = (long long)n * kPLB + kGuard + y * kPW + 2 * j;
                uint4 mk[DGRAD ? 2 : 1][4];
                if (DGRAD) {
#pragma unroll
                    for (int ox = 0; ox < 2; ++ox)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            mk[ox][c] = (rowok && 2 * j + ox < a.h_out)
                                            ? __ldg(reinterpret_cast<const uint4*>(a.mask + (c * a.cs_mask + row0 + ox) * 8))
                                            : make_uint4(0, 0, 0, 0);
                }
                mbar_wait(tfull + acc, acc_phase);
                tc_fence_after();
#pragma unroll
                for (int ox = 0; ox < 2; ++ox) {
                    float v32[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 128 + oy * 64 + ox * 32, v32);
                    if (ox == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty + acc);
                    }
                    const int x = 2 * j + ox;
                    if (rowok && x < a.h_out) {
                        uint32_t packed[16];
                        if (!DGRAD) {
#pragma unroll
                            for (int k = 0; k < 16; ++k)
                                packed[k] = pack_bf16x2(fmaxf(v32[2 * k] + bias_r[DGRAD ? 0 : 2 * k], 0.f),
                                                        fmaxf(v32[2 * k + 1] + bias_r[DGRAD ? 0 : 2 * k + 1], 0.f));
                        } else {
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const uint32_t mw[4] = {mk[DGRAD ? ox : 0][c].x, mk[DGRAD ? ox : 0][c].y, mk[DGRAD ? ox : 0][c].z,
                                                        mk[DGRAD ? ox : 0][c].w};
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    packed[4 * c + k] = pack_bf16x2(bf16_lo(mw[k]) > 0.f ? v32[8 * c + 2 * k] : 0.f,
                                                                    bf16_hi(mw[k]) > 0.f ? v32[8 * c + 2 * k + 1] : 0.f);
                            }
                        }
                        if (!DGRAD && a.out_mode == 2) {
                            // TB feature matrix, channel-group-major feature order (conv_tc.cu): unit (c/8)*h*h + y*h + x
                            const long long hw = (long long)a.h_out * a.h_out;
                            const long long u0 = (long long)y * a.h_out + x;
                            const int fr = n < a.feat_half ? n : n - a.feat_half + a.feat_half_row;
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                *reinterpret_cast<uint4*>(a.out + ((((long long)(fr >> 7)) * a.feat_rpad + c * hw + u0) * DRQ_TB_ACT + (fr & 127)) * 8) =
                                    make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
                        } else if (!DGRAD && a.out_mode == 1) {
                            uint4* dst = reinterpret_cast<uint4*>(a.out + (((long long)n * a.h_out + y) * a.h_out + x) * 32);
#pragma unroll
                            for (int c = 0; c < 4; ++c) dst[c] = make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
                        } else {
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                *reinterpret_cast<uint4*>(a.out + (c * a.cs_out + row0 + ox) * 8) =
                                    make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
                        }
                    }
                    if (DGRAD && ox == 1 && rowok && j == pitch - 1) {
                        // columns beyond the valid width are exact zeros (read by the next layer's dgrad / wgrad windows)
                        for (int x2 = a.h_out; x2 < kPW; ++x2)
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                *reinterpret_cast<uint4*>(a.out + (c * a.cs_out + row0 - 2 * j + x2) * 8) = make_uint4(0, 0, 0, 0);
                    }
                }
                if (++acc == kAcc) { acc = 0; acc_phase ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kAcc * 128);
}

constexpr size_t kSmem = kWBytes + kStages * kStageBytes + (2 * kStages + 2 * kAcc) * 8 + 16 + 128;

static uint32_t magic(int d) { return (uint32_t)((1ull << 32) / (uint32_t)d) + 1u; }

}  // namespace c2

// Forward (dgrad = 0): in = the layer's input activation (valid h_out + 2), out per out_mode.  Data gradient (dgrad = 1):
// in = the gradient of the layer's output (valid h_out_layer), out = the gradient of its input (valid h_out_layer + 2),
// masked by the input activation.
int conv2x2_launch(bool dgrad, const __nv_bfloat16* in, long long cs_in, const __nv_bfloat16* w, const float* bias,
                   const __nv_bfloat16* mask, long long cs_mask, __nv_bfloat16* out, long long cs_out, int N, int h_layer_out,
                   int out_mode, long long feat_rpad, int feat_half, int feat_half_row, cudaStream_t stream) {
    using namespace c2;
    Args a{};
    a.in = in; a.cs_in = cs_in; a.w = w; a.bias = bias; a.mask = mask; a.cs_mask = cs_mask; a.out = out; a.cs_out = cs_out;
    a.n_images = N;
    const int hin = h_layer_out + 2;
    a.pitch = (hin + 1) / 2;
    a.h_in = dgrad ? h_layer_out : hin;
    a.h_out = dgrad ? hin : h_layer_out;
    const int nrow = dgrad ? a.pitch + 1 : a.pitch;       // one block row of padding: the window of the first / last valid row
    a.pl4 = a.pitch * nrow;
    if (a.pitch > kMaxPitch || (long long)N * a.pl4 + kWS >= (1ll << 23)) {
        set_error("conv2x2: bad dims N=%d hout=%d", N, h_layer_out);
        return DRQ_ERR_INVALID;
    }
    a.total_pos = N * a.pl4;
    a.total_tiles = (a.total_pos + kTile - 1) / kTile;
    a.m_pl4 = magic(a.pl4);
    a.m_pitch = magic(a.pitch);
    a.out_mode = out_mode;
    a.feat_rpad = feat_rpad; a.feat_half = feat_half; a.feat_half_row = feat_half_row;
    const int grid = a.total_tiles < sm_budget() ? a.total_tiles : sm_budget();
    if (dgrad) {
        if (int rc = ensure_smem((const void*)conv2x2_tc_kernel<true>, kSmem, "conv2x2_dgrad")) return rc;
        launch_k(conv2x2_tc_kernel<true>, grid, kThreads, kSmem, stream, a);
        return check_launch("conv2x2_tc_kernel<dgrad>");
    }
    if (int rc = ensure_smem((const void*)conv2x2_tc_kernel<false>, kSmem, "conv2x2_fwd")) return rc;
    launch_k(conv2x2_tc_kernel<false>, grid, kThreads, kSmem, stream, a);
    return check_launch("conv2x2_tc_kernel<fwd>");
}

}  // namespace drq
