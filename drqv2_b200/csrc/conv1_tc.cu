// conv1 of the encoder (drqv2.py:55: Cin->32, k3, stride 2) on tensor cores, with
// RandomShiftsAug (integer shift, drqv2.py:19-45) and obs/255-0.5 (drqv2.py:64) fused into
// the loader: builder warps gather the augmented 3x3xCin patch of every output position
// straight from the uint8 frame stack into an im2col tile in shared memory (canonical
// no-swizzle UMMA layout [K unit][position][16 B]); the augmented image never exists in HBM.
//
//   forward : out[p][co] = relu(b'[co] + (1/255) sum_k im2col[p][k] * W[co][k])   M=128 pos, N=32, K=96
//   wgrad   : S[co][k]   = sum_p d[p][co] * im2col[p][k]                          M=64(co), N=96, K=pos
//
// The im2col entries are the exact integers x - 128 in bf16 (|x - 128| <= 128 has at most 8 significant
// bits, so no rounding of the pixels): a byte permute with 0x4B000000 turns a uint8 into the float
// 2^23 + x, one subtract centres it, one cvt packs two of them.  The affine map x/255 - 0.5 = (x - 128)/255 + (128/255 - 0.5) moves to the epilogues:
//   forward : b'[co] = b[co] + (128/255 - 0.5) sum_k bf16(W[co][k])   (pack kernel), accumulator scaled by 1/255
//   wgrad   : dW[co][k] = S[co][k]/255 + (128/255 - 0.5) db[co],  db[co] = S[co][cin*9]
// K = cin*9 is padded to 96; slot k = cin*9 holds the constant 1.0 so that column cin*9 of the
// weight-gradient GEMM is the bias gradient (the forward weight there is zero).  (kind::f16 needs both
// operands in the same 16-bit format - mixing an fp16 im2col with the bf16 gradient tile traps - hence
// bf16 rather than the cheaper fp16 construction.)
//
// Loader: the (vertically clamped) source rows of a tile are staged per channel as replicate-padded
// rows [4 | 84 | 4] bytes, so the horizontal shift is an address offset and the three kx taps of one
// (channel, ky) are three consecutive bytes: two aligned 32-bit loads + a funnel shift per group of
// three K entries instead of a byte load and a table lookup per entry.
//
// Warp roles (704 threads): warps 0..15 re-pitch the rows + build im2col tiles (4 threads per output
// position), warp 16 issues UMMAs (and bulk-loads the gradient tile in wgrad), warps 17..20 run the
// epilogue, warp 21 bulk-copies the source rows of the next tiles (4 stages ahead).
#include "pack.cuh"
#include "wgrad_reduce.cuh"

namespace drq {

using namespace tc;

constexpr int kC1K = 96, kC1Units = kC1K / 8;
constexpr int kC1Tile = 128;
constexpr int kC1ABytes = kC1Units * kC1Tile * 16;     // 24576
constexpr int kC1Stages = 3;
constexpr int kC1WBytes = kC1Units * 32 * 16;          // weights [12][32][16 B] (+ 32 fused biases behind them)
constexpr int kC1Builders = 512;                        // 16 builder warps: 4 threads per output position
constexpr int kC1Threads = kC1Builders + 6 * 32;       // + UMMA warp, 4 epilogue warps, row producer warp
constexpr int kC1DBytes = 4 * kC1Tile * 16;            // wgrad: d tile [4 blocks][128][16 B]
constexpr int kC1Acc = 4;
constexpr float kC1Scale = 1.0f / 255.0f;
constexpr float kC1Shift = 128.0f / 255.0f - 0.5f;     // (x - 128)/255 + kC1Shift == x/255 - 0.5

struct Conv1TcArgs {
    const uint8_t* obs; const int* shift; int cin, pad;
    const __nv_bfloat16* w;       // fwd: packed bf16 [12][32][8], then float bias'[32]
    __nv_bfloat16* out; long long cs_out;          // fwd: WB output
    const __nv_bfloat16* d; long long cs_d;        // wgrad: WB gradient (N images)
    float* partial;                                // wgrad: [grid][32][96]
    int n_images;
    long long* stamps;                             // debug: builder cycle totals of block 0 (or null)
    // rows straight from the replay ring (ring_frames != null; obs unused): see drq_conv1_fwd_bf16_ring
    const uint8_t* ring_frames; const int* ring_ep_start; const int* ring_idx;
    long long ring_capacity; int ring_B, ring_nstep, ring_stack, ring_frame_c;
};

// first byte of channel `c` of image `n`: the gathered stack, or the ring frame the stack is made of
// (replay_buffer.py:151,153 rows idx-1 / idx+nstep-1; dmc.py:98-109 stack of the last frames, reset frame repeated)
__device__ __forceinline__ const uint8_t* channel_plane(const Conv1TcArgs& a, int n, int c) {
    if (!a.ring_frames) return a.obs + ((long long)n * a.cin + c) * (kImg * kImg);
    const int j = c / a.ring_frame_c, cc = c - j * a.ring_frame_c;
    const bool is_next = n >= a.ring_B;
    const int b = is_next ? n - a.ring_B : n;
    const int t = is_next ? a.ring_idx[b] + a.ring_nstep - 1 : a.ring_idx[b] - 1;
    int r = t - (a.ring_stack - 1 - j);
    r = r < 0 ? 0 : r;
    const long long slot = ((long long)a.ring_ep_start[b] + r) % a.ring_capacity;
    return a.ring_frames + (slot * a.ring_frame_c + cc) * (long long)(kImg * kImg);
}

static long long* g_c1_stamps = nullptr;
#ifdef DRQ_STAMPS
#define C1_T() (a.stamps ? clock64() : 0ll)
#else
#define C1_T() 0ll
#endif

// Input rows of one tile staged in shared memory: per channel up to 11 (clamped) source rows, each as a
// replicate-padded row of kC1RowBytes.
constexpr int kC1RowWords = kImg / 4;                 // 21 source words per row
constexpr int kC1MaxRows = 11;
constexpr int kC1RowBytes = 96;                       // [4 pad | 84 | 4 pad | 4 slack]
constexpr int kC1ChBytes = kC1MaxRows * kC1RowBytes;  // 1056
constexpr int kC1ChWords = kC1MaxRows * kC1RowWords;  // 231 source words per channel
constexpr int kC1InBytes = 10 * kC1ChBytes;           // one padded-row buffer (cin <= 10)
constexpr int kC1RawSlot = 960;                       // raw rows of one channel: <= 15 + 11 * 84 bytes, 16-byte granular
constexpr int kC1RawBytes = 10 * kC1RawSlot;          // one raw stage
constexpr int kC1RawStages = 4;

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

// The row producer warp lands the tile's (vertically clamped) source rows of every channel in shared
// memory with one bulk-async copy per channel (the rows of a channel are contiguous in the frame stack);
// the builders then re-pitch them as replicate-padded rows (drqv2.py:26 F.pad(mode='replicate'),
// horizontally).  Thread (channel ci, word jw) walks the rows: no index arithmetic in the loop.
__device__ __forceinline__ void pad_rows(const uint8_t* __restrict__ raw, uint8_t* __restrict__ stage, int off0, int nrows,
                                         int cin, int b) {
    constexpr int RW = kC1RowBytes / 4;
    if (b < cin * kC1RowWords) {
        // thread (channel ci, word jw): copy the 84 pixel bytes of every row to byte 4 of the padded row
        const int ci = b / kC1RowWords, jw = b - ci * kC1RowWords;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(raw + ci * kC1RawSlot + off0) + jw;
        uint32_t* dst = reinterpret_cast<uint32_t*>(stage + ci * kC1ChBytes) + 1 + jw;
#pragma unroll 4
        for (int r = 0; r < nrows; ++r) dst[r * RW] = src[r * kC1RowWords];
    } else if (b < cin * kC1RowWords + 2 * cin) {
        // thread (channel ci, side): the replicate pad words - pixel 0 x4 on the left, pixel 83 x4 on the right
        const int e = b - cin * kC1RowWords, ci = e >> 1, side = e & 1;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(raw + ci * kC1RawSlot + off0) + (side ? kC1RowWords - 1 : 0);
        uint32_t* dst = reinterpret_cast<uint32_t*>(stage + ci * kC1ChBytes) + (side ? 1 + kC1RowWords : 0);
        const uint32_t sel = side ? 0x3333u : 0x0000u;
#pragma unroll 4
        for (int r = 0; r < nrows; ++r) dst[r * RW] = __byte_perm(src[r * kC1RowWords], 0, sel);
    }
}

// bf16 pair (x_A - 128, x_B - 128) from byte bA of group word gA (low half) and byte bB of gB (high half):
// a byte permute with 0x4B000000 makes the float 2^23 + x, subtracting 2^23 + 128 is exact, and so is the
// conversion to bf16 (|x - 128| <= 128 has at most 8 significant bits).
__device__ __forceinline__ float centred(uint32_t g, int b) {
    return __uint_as_float(__byte_perm(g, 0x4B000000u, (uint32_t)b | 0x7650u)) - 8388736.0f;
}
__device__ __forceinline__ uint32_t pair_from(uint32_t gA, int bA, uint32_t gB, int bB) {
    return pack_bf16x2(centred(gA, bA), centred(gB, bB));
}
__device__ __forceinline__ uint32_t one_from(uint32_t gA, int bA) {              // low half x - 128, high half 0
    return pack_bf16x2(centred(gA, bA), 0.f);
}

// One builder thread: position r of the tile, K units [U0, U0 + 3).  K entry k = 9*c + 3*ky + kx; group
// i = k / 3 = 3*c + ky is three consecutive bytes of channel c's staged row ky at byte offset g.
template <int U0, int CIN>                             // CIN = compile-time channel count, 0 = use cin
__device__ __forceinline__ void build_part(uint8_t* tile, const uint8_t* __restrict__ stage, int cin, int r,
                                           const int (&rowoff)[3], int g) {
    constexpr int G0 = U0 * 8 / 3;                     // first group of this quarter (0, 8, 16, 24)
    constexpr int NG = 8;                              // 24 K entries = 8 groups
    const int kmax = CIN ? CIN * 9 : cin * 9;
    uint32_t grp[NG];
    const int sh = (g & 3) * 8;
    const uint8_t* base[3];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) base[ky] = stage + rowoff[ky] + (g & ~3);
#pragma unroll
    for (int i = 0; i < NG; ++i) {
        const int gi = G0 + i, c = gi / 3, ky = gi % 3;
        uint32_t v = 0;
        if (gi * 3 < kmax) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(base[ky] + c * kC1ChBytes);
            v = __funnelshift_r(p[0], p[1], sh);
        }
        grp[i] = v;
    }
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int kA = (U0 + u) * 8 + 2 * e, kB = kA + 1;
            const int iA = kA / 3 - G0, iB = kB / 3 - G0;
            uint32_t v = 0;
            if (kB < kmax) v = pair_from(grp[iA], kA % 3, grp[iB], kB % 3);
            else if (kA < kmax) v = one_from(grp[iA], kA % 3) | (kB == kmax ? 0x3F800000u : 0u);   // data | 1.0
            else if (kA == kmax) v = 0x00003F80u;                                                    // 1.0 | 0
            w[e] = v;
        }
        *reinterpret_cast<uint4*>(tile + (U0 + u) * (kC1Tile * 16) + r * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

template <bool WGRAD, int CIN>
__global__ void __launch_bounds__(kC1Threads, 1) conv1_tc_kernel(Conv1TcArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int STAGE = kC1ABytes + (WGRAD ? kC1DBytes : 0);
    uint8_t* w_s = smem;                                    // fwd only: packed weights
    float* bias_s = reinterpret_cast<float*>(smem + kC1WBytes);   // fwd only: fused bias'
    uint8_t* in_s = smem + kC1WBytes + 128;                 // 2 builder groups x 2 padded-row buffers
    uint8_t* raw_s = in_s + 4 * kC1InBytes;                 // kC1RawStages x raw rows (bulk-copy destinations)
    uint8_t* st_s = raw_s + kC1RawStages * kC1RawBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(st_s + kC1Stages * STAGE + (WGRAD ? kC1DBytes : 0));
    uint64_t* full = bars;                     // one builder group (256 arrivals)
    uint64_t* empty = bars + kC1Stages;
    uint64_t* dfull = bars + 2 * kC1Stages;    // wgrad: bulk copy of the d tile
    uint64_t* tfull = bars + 3 * kC1Stages;
    uint64_t* tempty = bars + 3 * kC1Stages + kC1Acc;
    uint64_t* done = bars + 3 * kC1Stages + 2 * kC1Acc;
    uint64_t* rfull = done + 1;                // raw rows landed (cin producer lanes announce their bytes)
    uint64_t* rempty = rfull + kC1RawStages;   // raw stage re-pitched by the 256 builders of a group
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rempty + kC1RawStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int TILES_PER_IMG = (kPW * kPW + kC1Tile - 1) / kC1Tile;   // 14
    const int total_tiles = a.n_images * TILES_PER_IMG;

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int i = 0; i < kC1Stages; ++i) { mbar_init(full + i, kC1Builders / 2); mbar_init(empty + i, 1); mbar_init(dfull + i, 1); }
        for (int i = 0; i < kC1Acc; ++i) { mbar_init(tfull + i, 1); mbar_init(tempty + i, 4); }
        mbar_init(done, 1);
        for (int i = 0; i < kC1RawStages; ++i) { mbar_init(rfull + i, a.cin); mbar_init(rempty + i, kC1Builders / 2); }
        fence_barrier_init();
    }
    if (warp == 16) {
        tmem_alloc(tmem_slot, 128);
        tmem_relinquish();
    }
    pdl_wait();                 // CTA-local set-up above overlaps the previous kernel's tail
    if (!WGRAD) {
        const uint4* src = reinterpret_cast<const uint4*>(a.w);
        uint4* dst = reinterpret_cast<uint4*>(w_s);
        for (int i = threadIdx.x; i < (kC1WBytes + 128) / 16; i += kC1Threads) dst[i] = __ldg(src + i);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 16) {
        // ------------------------------------------------ row re-pitch + im2col builders
        // Two groups of 8 warps work on alternate tiles (group 0: tiles 0, 2, 4, ... of this CTA; group 1: 1, 3, ...)
        // so that one group's re-pitch / barrier latency overlaps the other's build.  Ring positions follow
        // from the tile ordinal.
        const int group = warp >> 3, gb = threadIdx.x & 255, r = gb & 127, half = gb >> 7;
        int it = 0;
        for (int ord = group; blockIdx.x + ord * (int)gridDim.x < total_tiles; ord += 2, ++it) {
            const int t = blockIdx.x + ord * gridDim.x;
            const int stage = ord % kC1Stages; const uint32_t phase = (ord / kC1Stages) & 1;
            const int rs = ord % kC1RawStages; const uint32_t rphase = (ord / kC1RawStages) & 1;
            const TileGeom cur = tile_geom(t, a.shift, a.pad);
            uint8_t* stg = in_s + (group * 2 + (it & 1)) * kC1InBytes;
            mbar_wait(rfull + rs, rphase);
            pad_rows(raw_s + rs * kC1RawBytes, stg, (cur.rlo * kImg) & 15, cur.nrows, a.cin, gb);
            mbar_arrive(rempty + rs);
            if (group == 0) asm volatile("bar.sync 1, 256;" ::: "memory");      // padded rows visible to the group
            else            asm volatile("bar.sync 2, 256;" ::: "memory");
            // Positions past the image (last tile) build from clamped, valid addresses in the forward (their rows
            // are never stored); in wgrad their im2col rows are zeroed below (the gradient tile's rows there
            // belong to the next image).
            const int p = min(cur.p0 + r, kPW * kPW - 1);
            const bool live = !WGRAD || cur.p0 + r < kPW * kPW;
            const int oy = p / kPW, ox = p - oy * kPW;
            const int sx = a.shift ? a.shift[2 * cur.n] : a.pad, sy = a.shift ? a.shift[2 * cur.n + 1] : a.pad;
            int rowoff[3];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
                rowoff[ky] = (clampi(2 * oy + ky + sy - a.pad, 0, kImg - 1) - cur.rlo) * kC1RowBytes;
            // aug column j = src[clamp(j + sx - pad)] = padded[j + sx + (4 - pad)]; the three kx taps start at j = 2*ox
            const int gofs = 2 * ox + sx + (4 - a.pad);
            mbar_wait(empty + stage, phase ^ 1);
            uint8_t* tile = st_s + stage * STAGE;
            if (!live) {
#pragma unroll
                for (int u = 0; u < 6; ++u)
                    *reinterpret_cast<uint4*>(tile + (half * 6 + u) * (kC1Tile * 16) + r * 16) = make_uint4(0, 0, 0, 0);
            } else if (half == 0) {
                build_part<0, CIN>(tile, stg, a.cin, r, rowoff, gofs);
                build_part<3, CIN>(tile, stg, a.cin, r, rowoff, gofs);
            } else {
                build_part<6, CIN>(tile, stg, a.cin, r, rowoff, gofs);
                build_part<9, CIN>(tile, stg, a.cin, r, rowoff, gofs);
            }
            fence_proxy_async();
            mbar_arrive(full + stage);
        }
    } else if (warp == 16) {
        // ------------------------------------------------ UMMA issuer (+ d-tile producer in wgrad)
        int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
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
        // descriptors: per-stage base + compile-time (address >> 4) offsets
        const uint64_t da0 = WGRAD ? make_smem_desc(smem_u32(st_s) + kC1ABytes, 128, kC1Tile * 16)     // d^T, MN-major
                                   : make_smem_desc(smem_u32(st_s), kC1Tile * 16, 128);               // im2col, K-major
        const uint64_t db0 = WGRAD ? make_smem_desc(smem_u32(st_s), 128, kC1Tile * 16)                 // im2col, MN-major
                                   : make_smem_desc(smem_u32(w_s), 512, 128);                         // weights, K-major
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            if (!WGRAD) mbar_wait(tempty + acc, acc_phase ^ 1);
            mbar_wait(full + stage, phase);
            if (WGRAD) mbar_wait(dfull + stage, phase);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t so = (uint64_t)(stage * (STAGE >> 4));
                if (!WGRAD) {
                    constexpr uint32_t idesc = make_idesc_bf16(128, 32, false, false);
#pragma unroll
                    for (int ks = 0; ks < kC1K / 16; ++ks)
                        umma_bf16(tmem_base + acc * 32, da0 + so + (uint64_t)(ks * 2 * kC1Tile), db0 + (uint64_t)(ks * 2 * 32), idesc,
                                  ks ? 1u : 0u);
                    umma_commit(empty + stage);
                    umma_commit(tfull + acc);
                } else {
                    // S[co (64)][k (96)] += d^T[co][pos] * im2col[pos][k]; both operands MN-major, K = 16 positions
                    constexpr uint32_t idesc = make_idesc_bf16(64, kC1K, true, true);
#pragma unroll
                    for (int ks = 0; ks < kC1Tile / 16; ++ks)
                        umma_bf16(tmem_base, da0 + so + (uint64_t)(ks * 16), db0 + so + (uint64_t)(ks * 16), idesc,
                                  (first && ks == 0) ? 0u : 1u);
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
    } else if (warp == 21) {
        // ------------------------------------------------ row producer: lane c copies channel c's rows of the tile
        if (lane < a.cin) {
            int rs = 0; uint32_t rphase = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const TileGeom g = tile_geom(t, a.shift, a.pad);
                const uint8_t* src = channel_plane(a, g.n, lane) + g.rlo * kImg;
                const int off0 = (g.rlo * kImg) & 15;            // every channel plane starts 16-byte aligned
                const uint32_t nbytes = (uint32_t)((off0 + g.nrows * kImg + 15) & ~15);
                mbar_wait(rempty + rs, rphase ^ 1);
                mbar_arrive_expect_tx(rfull + rs, nbytes);
                bulk_g2s(raw_s + rs * kC1RawBytes + lane * kC1RawSlot, src - off0, nbytes, rfull + rs);
                if (++rs == kC1RawStages) { rs = 0; rphase ^= 1; }
            }
        }
        pdl_release();                          // last rows are on their way: the next kernel may set itself up
    } else {
        // ------------------------------------------------ epilogue warps 17..20 -> lane quarters 1,2,3,0
        const int q = warp & 3;
        if (!WGRAD) {
            int acc = 0; uint32_t acc_phase = 0;
            float bias_r[32];                       // registers: shared-memory reads per tile were bank-conflict wavefronts
#pragma unroll
            for (int i = 0; i < 32; ++i) bias_r[i] = bias_s[i];
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
                    packed[i] = pack_bf16x2(fmaxf(fmaf(v[2 * i], kC1Scale, bias_r[2 * i]), 0.f),
                                            fmaxf(fmaf(v[2 * i + 1], kC1Scale, bias_r[2 * i + 1]), 0.f));
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
    if (warp == 16) tmem_dealloc(tmem_base, 128);
}

__global__ void __launch_bounds__(256)
pack_conv1_w_kernel(const float* __restrict__ w, const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int cin) {
    pdl_trigger();
    pdl_wait();
    __shared__ float part[32][9];
    pack_conv1_w_block(w, bias, out, cin, threadIdx.x, part);
}

// dw[co][k] = S[co][k]/255 + (128/255 - 0.5) db[co] (k < cin*9), db[co] = S[co][cin*9], S = sum of the per-CTA
// partials [32][96].  Block = 32 outputs x 8 slices of the G partials, combined in fixed order.
__global__ void __launch_bounds__(256)
conv1_wgrad_reduce_kernel(const float* __restrict__ partial, int G, int cin, float* __restrict__ dw, float* __restrict__ db) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[8][33], redb[8][33];
    conv1_reduce_block(partial, G, cin, dw, db, blockIdx.x, red, redb);
}

constexpr size_t kConv1FwdSmem = kC1WBytes + 128 + 4 * kC1InBytes + kC1RawStages * kC1RawBytes + kC1Stages * kC1ABytes +
                                 (3 * kC1Stages + 2 * kC1Acc + 1 + 2 * kC1RawStages) * 8 + 16;
// wgrad: + one extra d-tile worth of tail padding (rows 32..63 of the M=64 operand read 4 blocks past the tile)
constexpr size_t kConv1WgSmem = kC1WBytes + 128 + 4 * kC1InBytes + kC1RawStages * kC1RawBytes + kC1Stages * (kC1ABytes + kC1DBytes) +
                                kC1DBytes + (3 * kC1Stages + 2 * kC1Acc + 1 + 2 * kC1RawStages) * 8 + 16;

}  // namespace drq

using namespace drq;

extern "C" {

int drq_debug_conv1_stamps(int64_t* buf) { g_c1_stamps = reinterpret_cast<long long*>(buf); return DRQ_OK; }

int64_t drq_conv1_w_packed_elems(void) { return kC1Units * 32 * 8 + 64; }

int drq_pack_conv1_w_bf16(const float* w, const float* bias, uint16_t* out, int cin, void* stream) {
    DRQ_REQUIRE(w && bias && out && cin > 0 && cin * 9 + 1 <= kC1K, "pack_conv1_w: bad args (cin <= 10)");
    launch_k(pack_conv1_w_kernel, 1, 256, 0, as_stream(stream), w, bias, reinterpret_cast<__nv_bfloat16*>(out), cin);
    return check_launch("pack_conv1_w_kernel");
}

// the ring view of a batch as kernel arguments
static int ring_args(Conv1TcArgs& a, const drq_ring_src* src, int B, int N) {
    DRQ_REQUIRE(src && src->frames && src->ep_start && src->idx, "conv1_*_ring: incomplete ring source");
    DRQ_REQUIRE(src->capacity > 0 && src->frame_c > 0 && src->stack > 0 && src->nstep > 0, "conv1_*_ring: bad ring dims");
    DRQ_REQUIRE(B > 0 && N > 0 && N <= 2 * B, "conv1_*_ring: N images must be B (obs) or 2B (obs | next_obs)");
    DRQ_REQUIRE(((uintptr_t)src->frames % 16) == 0, "conv1_*_ring: ring frames must be 16-byte aligned");
    a.ring_frames = src->frames; a.ring_ep_start = src->ep_start; a.ring_idx = src->idx;
    a.ring_capacity = src->capacity; a.ring_B = B; a.ring_nstep = src->nstep; a.ring_stack = src->stack;
    a.ring_frame_c = src->frame_c;
    a.cin = src->frame_c * src->stack;
    return DRQ_OK;
}

static int conv1_fwd_launch(Conv1TcArgs& a, const int32_t* shift, const uint16_t* w_packed, uint16_t* out, int N, int pad,
                            cudaStream_t stream);
static int conv1_wgrad_launch(Conv1TcArgs& a, const int32_t* shift, const uint16_t* dpre, float* partial, float* dw, float* db,
                              int N, int pad, cudaStream_t stream);

int drq_conv1_fwd_bf16_ring(const drq_ring_src* src, int B, const int32_t* shift, const uint16_t* w_packed,
                            uint16_t* out, int N, int pad, void* stream) {
    Conv1TcArgs a{};
    if (int rc = ring_args(a, src, B, N)) return rc;
    return conv1_fwd_launch(a, shift, w_packed, out, N, pad, as_stream(stream));
}

int drq_conv1_wgrad_bf16_ring(const drq_ring_src* src, int B, const int32_t* shift, const uint16_t* dpre,
                              float* partial, float* dw, float* db, int N, int pad, void* stream) {
    Conv1TcArgs a{};
    if (int rc = ring_args(a, src, B, N)) return rc;
    return conv1_wgrad_launch(a, shift, dpre, partial, dw, db, N, pad, as_stream(stream));
}

int drq_conv1_fwd_bf16(const uint8_t* obs, const int32_t* shift, const uint16_t* w_packed, uint16_t* out, int N,
                       int cin, int pad, void* stream) {
    DRQ_REQUIRE(obs, "conv1_fwd_bf16: null pointer");
    Conv1TcArgs a{};
    a.obs = obs; a.cin = cin;
    return conv1_fwd_launch(a, shift, w_packed, out, N, pad, as_stream(stream));
}

static int conv1_fwd_launch(Conv1TcArgs& a, const int32_t* shift, const uint16_t* w_packed, uint16_t* out, int N, int pad,
                            cudaStream_t stream) {
    const int cin = a.cin;
    DRQ_REQUIRE(w_packed && out, "conv1_fwd_bf16: null pointer");
    DRQ_REQUIRE(N > 0 && cin > 0 && cin * 9 + 1 <= kC1K && pad >= 0 && pad <= 4, "conv1_fwd_bf16: bad dims (cin <= 10, pad <= 4)");
    if (int rc = ensure_smem((const void*)conv1_tc_kernel<false, 9>, kConv1FwdSmem, "conv1_fwd_bf16")) return rc;
    if (int rc = ensure_smem((const void*)conv1_tc_kernel<false, 0>, kConv1FwdSmem, "conv1_fwd_bf16")) return rc;
    a.shift = shift; a.pad = pad;
    a.w = reinterpret_cast<const __nv_bfloat16*>(w_packed);
    a.out = reinterpret_cast<__nv_bfloat16*>(out);
    a.cs_out = (long long)N * DRQ_PLB + DRQ_WB_SLACK;
    a.n_images = N;
    a.stamps = g_c1_stamps;
    const int tiles = N * 14;
    const int G = tiles < sm_budget() ? tiles : sm_budget();
    if (cin == 9) launch_k(conv1_tc_kernel<false, 9>, G, kC1Threads, kConv1FwdSmem, stream, a);
    else launch_k(conv1_tc_kernel<false, 0>, G, kC1Threads, kConv1FwdSmem, stream, a);
    return check_launch("conv1_tc_kernel<fwd>");
}

int64_t drq_conv1_wgrad_bf16_ws_floats(void) { return 148ll * 32 * kC1K; }

int drq_conv1_wgrad_bf16(const uint8_t* obs, const int32_t* shift, const uint16_t* dpre, float* partial,
                         float* dw, float* db, int N, int cin, int pad, void* stream) {
    DRQ_REQUIRE(obs, "conv1_wgrad_bf16: null pointer");
    Conv1TcArgs a{};
    a.obs = obs; a.cin = cin;
    return conv1_wgrad_launch(a, shift, dpre, partial, dw, db, N, pad, as_stream(stream));
}

static int conv1_wgrad_launch(Conv1TcArgs& a, const int32_t* shift, const uint16_t* dpre, float* partial, float* dw, float* db,
                              int N, int pad, cudaStream_t stream) {
    const int cin = a.cin;
    DRQ_REQUIRE(dpre && partial && (dw != nullptr) == (db != nullptr), "conv1_wgrad_bf16: null pointer");
    DRQ_REQUIRE(N > 0 && cin > 0 && cin * 9 + 1 <= kC1K && pad >= 0 && pad <= 4, "conv1_wgrad_bf16: bad dims (cin <= 10, pad <= 4)");
    if (int rc = ensure_smem((const void*)conv1_tc_kernel<true, 9>, kConv1WgSmem, "conv1_wgrad_bf16")) return rc;
    if (int rc = ensure_smem((const void*)conv1_tc_kernel<true, 0>, kConv1WgSmem, "conv1_wgrad_bf16")) return rc;
    a.shift = shift; a.pad = pad;
    a.d = reinterpret_cast<const __nv_bfloat16*>(dpre);
    a.cs_d = (long long)N * DRQ_PLB + DRQ_WB_SLACK;
    a.partial = partial;
    a.n_images = N;
    const int G = conv1_wgrad_ctas(N);
    if (cin == 9) launch_k(conv1_tc_kernel<true, 9>, G, kC1Threads, kConv1WgSmem, stream, a);
    else launch_k(conv1_tc_kernel<true, 0>, G, kC1Threads, kConv1WgSmem, stream, a);
    if (int rc = check_launch("conv1_tc_kernel<wgrad>")) return rc;
    if (!dw) return DRQ_OK;                                   // partials only: reduced later by drq_conv_wgrad_reduce_multi
    launch_k(conv1_wgrad_reduce_kernel, kC1ReduceBlocks, 256, 0, stream, partial, G, cin, dw, db);
    return check_launch("conv1_wgrad_reduce_kernel");
}

}  // extern "C"
