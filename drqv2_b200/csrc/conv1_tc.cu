// conv1 of the encoder (drqv2.py:55: Cin->32, k3, stride 2) on tensor cores, with
// RandomShiftsAug (integer shift, drqv2.py:19-45) and obs/255-0.5 (drqv2.py:64) fused into
// the loader: builder warps gather the augmented, normalised 3x3xCin patch of every output
// position straight from the uint8 frame stack into an im2col tile in shared memory (canonical
// no-swizzle UMMA layout [K unit][position][16 B]); the augmented image never exists in HBM.
//
//   forward : out[p][co] = relu(b[co] + sum_k im2col[p][k] * W[co][k])     M=128 pos, N=32, K=96
//   wgrad   : dW[co][k]  = sum_p d[p][co] * im2col[p][k]                   M=64(co), N=96, K=pos
// K = cin*9 is padded to 96; slot k = cin*9 holds the constant 1.0 so that the weight-gradient
// GEMM's column cin*9 is the bias gradient (the forward weight there is zero).
//
// Warp roles (416 threads): warps 0..7 build im2col tiles, warp 8 issues UMMAs (and bulk-loads
// the gradient tile in wgrad), warps 9..12 run the epilogue.
#include "tc_common.cuh"

namespace drq {

using namespace tc;

constexpr int kC1K = 96, kC1Units = kC1K / 8;
constexpr int kC1Tile = 128;
constexpr int kC1ABytes = kC1Units * kC1Tile * 16;     // 24576
constexpr int kC1Stages = 3;
constexpr int kC1WBytes = kC1Units * 32 * 16;          // weights [12][32][16 B]
constexpr int kC1Threads = 13 * 32;
constexpr int kC1DBytes = 4 * kC1Tile * 16;            // wgrad: d tile [4 blocks][128][16 B]
constexpr int kC1Acc = 4;

struct Conv1TcArgs {
    const uint8_t* obs; const int* shift; int cin, pad;
    const __nv_bfloat16* w;       // fwd: packed [12][32][8]
    const float* bias;
    __nv_bfloat16* out; long long cs_out;          // fwd: WB output
    const __nv_bfloat16* d; long long cs_d;        // wgrad: WB gradient (N images)
    float* partial;                                // wgrad: [grid][32][96]
    int n_images;
};

// Input rows of one tile staged in shared memory: per channel the contiguous byte range of the
// (clamped) source rows [rlo, rhi], at most 11 rows x 84 B = 231 words.
constexpr int kC1RowWords = kImg / 4;                 // 21
constexpr int kC1ChWords = 11 * kC1RowWords;          // 231
constexpr int kC1ChBytes = 928;                       // 231 words padded to 16 B
constexpr int kC1StageWords = 10;                     // ceil(10 ch * 231 / 256) words per builder thread (cin <= 10)
constexpr int kC1InBytes = 10 * kC1ChBytes;           // one staging buffer

struct TileGeom { int n, p0, rlo, nrows; };

__device__ __forceinline__ TileGeom tile_geom(int t, const int* __restrict__ shift, int pad) {
    constexpr int TILES_PER_IMG = (kPW * kPW + kC1Tile - 1) / kC1Tile;
    TileGeom g;
    g.n = t / TILES_PER_IMG;
    g.p0 = (t - g.n * TILES_PER_IMG) * kC1Tile;
    const int sy = shift ? shift[2 * g.n + 1] : pad;
    const int oy_min = g.p0 / kPW;
    const int oy_max = min(g.p0 + kC1Tile - 1, kPW * kPW - 1) / kPW;
    g.rlo = clampi(2 * oy_min + sy - pad, 0, kImg - 1);
    const int rhi = clampi(2 * oy_max + 2 + sy - pad, 0, kImg - 1);
    g.nrows = rhi - g.rlo + 1;
    return g;
}

// each builder thread fetches up to kC1StageWords 4-byte words of the tile's input rows
__device__ __forceinline__ void fetch_rows(uint32_t (&regs)[kC1StageWords], const uint8_t* __restrict__ obs,
                                           const TileGeom& g, int cin, int b) {
    const uint32_t* img = reinterpret_cast<const uint32_t*>(obs + (long long)g.n * cin * kImg * kImg);
    const int nwords = g.nrows * kC1RowWords;
#pragma unroll
    for (int i = 0; i < kC1StageWords; ++i) {
        const int w = b + 256 * i;
        const int ci = w / kC1ChWords, j = w - ci * kC1ChWords;
        regs[i] = (ci < cin && j < nwords) ? __ldg(img + ci * (kImg * kImg / 4) + g.rlo * kC1RowWords + j) : 0u;
    }
}
__device__ __forceinline__ void store_rows(const uint32_t (&regs)[kC1StageWords], uint8_t* stage, int b) {
#pragma unroll
    for (int i = 0; i < kC1StageWords; ++i) {
        const int w = b + 256 * i;
        const int ci = w / kC1ChWords, j = w - ci * kC1ChWords;
        if (ci < 10) *reinterpret_cast<uint32_t*>(stage + ci * kC1ChBytes + 4 * j) = regs[i];
    }
}

// one builder thread: position r of the tile, K range [K0, K0+48); pixels come from the staged
// rows, the u8 -> bf16(x/255 - 0.5) map (drqv2.py:64) from a 256-entry table
template <int K0>
__device__ __forceinline__ void build_half(uint8_t* tile, const uint8_t* __restrict__ stage,
                                           const uint16_t* __restrict__ lut, int cin, int r, bool valid,
                                           const int (&off9)[9]) {
    const int kmax = cin * 9;
#pragma unroll
    for (int u = 0; u < 6; ++u) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            uint32_t hv[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = K0 + u * 8 + e * 2 + h;
                const int ci = k / 9, t = k % 9;
                uint32_t v = 0;
                if (valid) {
                    if (k < kmax) v = lut[stage[ci * kC1ChBytes + off9[t]]];
                    else if (k == kmax) v = 0x3F80u;       // bf16 1.0: the bias-gradient column
                }
                hv[h] = v;
            }
            w[e] = hv[0] | (hv[1] << 16);
        }
        *reinterpret_cast<uint4*>(tile + (K0 / 8 + u) * (kC1Tile * 16) + r * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

template <bool WGRAD>
__global__ void __launch_bounds__(kC1Threads, 1) conv1_tc_kernel(Conv1TcArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int STAGE = kC1ABytes + (WGRAD ? kC1DBytes : 0);
    uint8_t* w_s = smem;                                    // fwd only
    uint8_t* in_s = smem + kC1WBytes;                       // 2 x staged input rows
    uint16_t* lut_s = reinterpret_cast<uint16_t*>(in_s + 2 * kC1InBytes);
    uint8_t* st_s = in_s + 2 * kC1InBytes + 512;
    uint64_t* bars = reinterpret_cast<uint64_t*>(st_s + kC1Stages * STAGE + (WGRAD ? kC1DBytes : 0));
    uint64_t* full = bars;                     // builders (256 arrivals) [+ d-tile tx in wgrad: separate barrier]
    uint64_t* empty = bars + kC1Stages;
    uint64_t* dfull = bars + 2 * kC1Stages;    // wgrad: bulk copy of the d tile
    uint64_t* tfull = bars + 3 * kC1Stages;
    uint64_t* tempty = bars + 3 * kC1Stages + kC1Acc;
    uint64_t* done = bars + 3 * kC1Stages + 2 * kC1Acc;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int TILES_PER_IMG = (kPW * kPW + kC1Tile - 1) / kC1Tile;   // 14
    const int total_tiles = a.n_images * TILES_PER_IMG;

    if (!WGRAD) {
        const uint4* src = reinterpret_cast<const uint4*>(a.w);
        uint4* dst = reinterpret_cast<uint4*>(w_s);
        for (int i = threadIdx.x; i < kC1WBytes / 16; i += kC1Threads) dst[i] = __ldg(src + i);
    }
    if (threadIdx.x < 256)
        lut_s[threadIdx.x] = __bfloat16_as_ushort(__float2bfloat16_rn(
            __fsub_rn(__fdiv_rn((float)threadIdx.x, 255.0f), 0.5f)));          // drqv2.py:64, then bf16
    if (threadIdx.x == 0) {
        for (int i = 0; i < kC1Stages; ++i) { mbar_init(full + i, 256); mbar_init(empty + i, 1); mbar_init(dfull + i, 1); }
        for (int i = 0; i < kC1Acc; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 4); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 8) {
        tmem_alloc(tmem_slot, 128);
        tmem_relinquish();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ------------------------------------------------ im2col builders
        const int b = threadIdx.x, r = b & 127, half = b >> 7;
        int stage = 0; uint32_t phase = 0; int buf = 0;
        uint32_t regs[kC1StageWords];
        TileGeom g = tile_geom(blockIdx.x, a.shift, a.pad);
        if ((int)blockIdx.x < total_tiles) fetch_rows(regs, a.obs, g, a.cin, b);
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            uint8_t* stg = in_s + buf * kC1InBytes;
            store_rows(regs, stg, b);
            asm volatile("bar.sync 1, 256;" ::: "memory");      // staged rows visible to all builders
            const TileGeom cur = g;
            const int tn = t + gridDim.x;
            if (tn < total_tiles) {                              // prefetch the next tile's rows (in flight during the build)
                g = tile_geom(tn, a.shift, a.pad);
                fetch_rows(regs, a.obs, g, a.cin, b);
            }
            const int p = cur.p0 + r;
            const bool valid = p < kPW * kPW;
            const int oy = p / kPW, ox = p - oy * kPW;
            const int sx = a.shift ? a.shift[2 * cur.n] : a.pad, sy = a.shift ? a.shift[2 * cur.n + 1] : a.pad;
            int off9[9];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int sr = clampi(2 * oy + ky + sy - a.pad, 0, kImg - 1) - cur.rlo;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
                    off9[ky * 3 + kx] = valid ? sr * kImg + clampi(2 * ox + kx + sx - a.pad, 0, kImg - 1) : 0;
            }
            mbar_wait(empty + stage, phase ^ 1);
            uint8_t* tile = st_s + stage * STAGE;
            if (half == 0) build_half<0>(tile, stg, lut_s, a.cin, r, valid, off9);
            else           build_half<48>(tile, stg, lut_s, a.cin, r, valid, off9);
            fence_proxy_async();
            mbar_arrive(full + stage);
            if (++stage == kC1Stages) { stage = 0; phase ^= 1; }
            buf ^= 1;
        }
    } else if (warp == 8) {
        // ------------------------------------------------ UMMA issuer (+ d-tile producer in wgrad)
        int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
        const uint32_t w_addr = smem_u32(w_s);
        bool first = true;
        if (WGRAD && elect_one()) {
            // prefetch the d tiles of the first kC1Stages tiles
            int s2 = 0;
            for (int t = blockIdx.x; t < total_tiles && s2 < kC1Stages; t += gridDim.x, ++s2) {
                const int n = t / TILES_PER_IMG, p0 = (t - n * TILES_PER_IMG) * kC1Tile;
                const long long row0 = (long long)n * DRQ_PLB + DRQ_GUARD + p0;
                mbar_arrive_expect_tx(dfull + s2, kC1DBytes);
                for (int c = 0; c < 4; ++c)
                    bulk_g2s(st_s + s2 * STAGE + kC1ABytes + c * kC1Tile * 16, a.d + (c * a.cs_d + row0) * 8, kC1Tile * 16, dfull + s2);
            }
        }
        __syncwarp();
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            if (!WGRAD) mbar_wait(tempty + acc, acc_phase ^ 1);
            mbar_wait(full + stage, phase);
            if (WGRAD) mbar_wait(dfull + stage, phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_addr = smem_u32(st_s + stage * STAGE);
                if (!WGRAD) {
                    constexpr uint32_t idesc = make_idesc_bf16(128, 32, false, false);
#pragma unroll
                    for (int ks = 0; ks < kC1K / 16; ++ks) {
                        const uint64_t da = make_smem_desc(a_addr + ks * 2 * kC1Tile * 16, kC1Tile * 16, 128);
                        const uint64_t db = make_smem_desc(w_addr + ks * 2 * 512, 512, 128);
                        umma_bf16(tmem_base + acc * 32, da, db, idesc, ks ? 1u : 0u);
                    }
                    umma_commit(empty + stage);
                    umma_commit(tfull + acc);
                } else {
                    // D[co (64)][k (96)] += d^T[co][pos] * im2col[pos][k]; both operands MN-major, K = 16 positions
                    constexpr uint32_t idesc = make_idesc_bf16(64, kC1K, true, true);
                    const uint32_t d_addr = a_addr + kC1ABytes;
#pragma unroll
                    for (int ks = 0; ks < kC1Tile / 16; ++ks) {
                        const uint64_t da = make_smem_desc(d_addr + ks * 256, 128, kC1Tile * 16);
                        const uint64_t db = make_smem_desc(a_addr + ks * 256, 128, kC1Tile * 16);
                        umma_bf16(tmem_base, da, db, idesc, (first && ks == 0) ? 0u : 1u);
                    }
                    umma_commit(empty + stage);
                }
            }
            __syncwarp();
            first = false;
            if (WGRAD) {
                // refill this stage's d tile for the tile kC1Stages ahead, once the UMMAs reading it retire
                const int tn = t + kC1Stages * gridDim.x;
                if (tn < total_tiles && elect_one()) {
                    mbar_wait(empty + stage, phase);
                    const int n = tn / TILES_PER_IMG, p0 = (tn - n * TILES_PER_IMG) * kC1Tile;
                    const long long row0 = (long long)n * DRQ_PLB + DRQ_GUARD + p0;
                    mbar_arrive_expect_tx(dfull + stage, kC1DBytes);
                    for (int c = 0; c < 4; ++c)
                        bulk_g2s(st_s + stage * STAGE + kC1ABytes + c * kC1Tile * 16, a.d + (c * a.cs_d + row0) * 8, kC1Tile * 16, dfull + stage);
                }
                __syncwarp();
            }
            if (++stage == kC1Stages) { stage = 0; phase ^= 1; }
            if (++acc == kC1Acc) { acc = 0; acc_phase ^= 1; }
        }
        if (WGRAD) {
            if (elect_one()) umma_commit(done);
            __syncwarp();
        }
    } else {
        // ------------------------------------------------ epilogue warps 9..12 -> lane quarters 1,2,3,0
        const int q = warp & 3;
        if (!WGRAD) {
            int acc = 0; uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int n = t / TILES_PER_IMG, p0 = (t - n * TILES_PER_IMG) * kC1Tile;
                mbar_wait(tfull + acc, acc_phase);
                tc_fence_after();
                float v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 32, v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty + acc);
                if (++acc == kC1Acc) { acc = 0; acc_phase ^= 1; }
                const int p = p0 + q * 32 + lane;
                if (p >= kPW * kPW) continue;
                const long long row = (long long)n * DRQ_PLB + DRQ_GUARD + p;
                uint32_t packed[16];
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    packed[i] = pack_bf16x2(fmaxf(v[2 * i] + __ldg(a.bias + 2 * i), 0.f),
                                            fmaxf(v[2 * i + 1] + __ldg(a.bias + 2 * i + 1), 0.f));
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<uint4*>(a.out + (c * a.cs_out + row) * 8) =
                        make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
            }
        } else {
            mbar_wait(done, 0);
            tc_fence_after();
            if (q < 2) {
                float* out = a.partial + (long long)blockIdx.x * (32 * kC1K);
#pragma unroll
                for (int c0 = 0; c0 < kC1K; c0 += 32) {
                    float v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
                    if (lane < 16) {
                        float4* dst = reinterpret_cast<float4*>(out + (q * 16 + lane) * kC1K + c0);
#pragma unroll
                        for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 128);
}

// fp32 conv1 weight [32][cin][3][3] -> bf16 [12 K units][32 co][8] (zero padded to K = 96)
__global__ void pack_conv1_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cin) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kC1Units * 32 * 8) return;
    const int u = i / 256, co = (i / 8) % 32, e = i % 8;
    const int k = u * 8 + e;
    out[i] = __float2bfloat16_rn(k < cin * 9 ? w[co * cin * 9 + k] : 0.f);
}

// dw[co][k] (k < cin*9) and db[co] (column cin*9) from per-CTA partials [32][96]
__global__ void conv1_wgrad_reduce_kernel(const float* __restrict__ partial, int G, int cin,
                                          float* __restrict__ dw, float* __restrict__ db) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 32 * kC1K) return;
    const int co = i / kC1K, k = i - co * kC1K;
    if (k > cin * 9) return;
    float s = 0.f;
    for (int g = 0; g < G; ++g) s += partial[(long long)g * (32 * kC1K) + i];
    if (k < cin * 9) dw[co * cin * 9 + k] = s; else db[co] = s;
}

constexpr size_t kConv1FwdSmem = kC1WBytes + 2 * kC1InBytes + 512 + kC1Stages * kC1ABytes + (3 * kC1Stages + 2 * kC1Acc + 1) * 8 + 16;
// wgrad: + one extra d-tile worth of tail padding (rows 32..63 of the M=64 operand read 4 blocks past the tile)
constexpr size_t kConv1WgSmem = kC1WBytes + 2 * kC1InBytes + 512 + kC1Stages * (kC1ABytes + kC1DBytes) + kC1DBytes +
                                (3 * kC1Stages + 2 * kC1Acc + 1) * 8 + 16;

}  // namespace drq

using namespace drq;

extern "C" {

int drq_pack_conv1_w_bf16(const float* w, uint16_t* out, int cin, void* stream) {
    DRQ_REQUIRE(w && out && cin > 0 && cin * 9 + 1 <= kC1K, "pack_conv1_w: bad args (cin <= 10)");
    pack_conv1_w_kernel<<<(kC1Units * 256 + 255) / 256, 256, 0, as_stream(stream)>>>(
        w, reinterpret_cast<__nv_bfloat16*>(out), cin);
    return check_launch("pack_conv1_w_kernel");
}

int drq_conv1_fwd_bf16(const uint8_t* obs, const int32_t* shift, const uint16_t* w_packed, const float* bias,
                       uint16_t* out, int N, int cin, int pad, void* stream) {
    DRQ_REQUIRE(obs && w_packed && bias && out, "conv1_fwd_bf16: null pointer");
    DRQ_REQUIRE(N > 0 && cin > 0 && cin * 9 + 1 <= kC1K && pad >= 0, "conv1_fwd_bf16: bad dims (cin <= 10)");
    if (int rc = ensure_smem((const void*)conv1_tc_kernel<false>, kConv1FwdSmem, "conv1_fwd_bf16")) return rc;
    Conv1TcArgs a{};
    a.obs = obs; a.shift = shift; a.cin = cin; a.pad = pad;
    a.w = reinterpret_cast<const __nv_bfloat16*>(w_packed);
    a.bias = bias;
    a.out = reinterpret_cast<__nv_bfloat16*>(out);
    a.cs_out = (long long)N * DRQ_PLB + DRQ_WB_SLACK;
    a.n_images = N;
    const int tiles = N * 14;
    conv1_tc_kernel<false><<<tiles < 148 ? tiles : 148, kC1Threads, kConv1FwdSmem, as_stream(stream)>>>(a);
    return check_launch("conv1_tc_kernel<fwd>");
}

int64_t drq_conv1_wgrad_bf16_ws_floats(void) { return 148ll * 32 * kC1K; }

int drq_conv1_wgrad_bf16(const uint8_t* obs, const int32_t* shift, const uint16_t* dpre, float* partial,
                         float* dw, float* db, int N, int cin, int pad, void* stream) {
    DRQ_REQUIRE(obs && dpre && partial && dw && db, "conv1_wgrad_bf16: null pointer");
    DRQ_REQUIRE(N > 0 && cin > 0 && cin * 9 + 1 <= kC1K && pad >= 0, "conv1_wgrad_bf16: bad dims (cin <= 10)");
    if (int rc = ensure_smem((const void*)conv1_tc_kernel<true>, kConv1WgSmem, "conv1_wgrad_bf16")) return rc;
    Conv1TcArgs a{};
    a.obs = obs; a.shift = shift; a.cin = cin; a.pad = pad;
    a.d = reinterpret_cast<const __nv_bfloat16*>(dpre);
    a.cs_d = (long long)N * DRQ_PLB + DRQ_WB_SLACK;
    a.partial = partial;
    a.n_images = N;
    const int tiles = N * 14;
    const int G = tiles < 148 ? tiles : 148;
    conv1_tc_kernel<true><<<G, kC1Threads, kConv1WgSmem, as_stream(stream)>>>(a);
    if (int rc = check_launch("conv1_tc_kernel<wgrad>")) return rc;
    conv1_wgrad_reduce_kernel<<<(32 * kC1K + 127) / 128, 128, 0, as_stream(stream)>>>(partial, G, cin, dw, db);
    return check_launch("conv1_wgrad_reduce_kernel");
}

}  // extern "C"
