// bf16 tensor-core GEMM (tcgen05 + TMEM + bulk-async copies) for the heads in bf16 mode: trunk
// Linear(39200->F) forward (split-K) / weight gradient / data gradient and the actor / twin-Q MLP
// layers.  Reference ops: nn.Linear forward/backward of drqv2.py:74-81,100-111.
//
//   C[M,N] = sum_k A(m,k) * B(n,k),  bf16 operands, fp32 accumulation in TMEM.
//
// Operands live in HBM in the feature-blocked layout "FB":  X_fb[f/8][row][8]  (16-byte units of
// 8 consecutive features, `rpad` rows per unit block, rows and features zero padded).  One
// buffer serves both roles a matrix plays in training:
//   * K-major  (contraction over the blocked feature dim):  unit block u holds K unit u of all rows;
//   * MN-major (contraction over the row dim, e.g. the batch in a weight gradient): unit block u
//     holds M/N unit u, rows are K.
// Either way a 128 x BK operand tile is a handful of contiguous 1-2 KB pieces, so one elected
// thread stages it with 1-D bulk-async copies (TMA engine, mbarrier complete_tx) directly in the
// canonical no-swizzle UMMA layout [unit][line][16 B] - no per-thread address math, no proxy
// fences, 8-deep pipeline.  Warp roles (192 threads): warp 0 = copy producer, warp 1 = UMMA
// issuer, warps 2..5 = epilogue (TMEM lane quarters 2,3,0,1).
#include "tc_common.cuh"

namespace drq {

using namespace tc;

constexpr int GT_BM = 128, GT_BK = 64, GT_THREADS = 192;

__host__ __device__ constexpr int gt_stages(int bn) { return bn >= 128 ? 6 : 8; }

struct GemmTcArgs {
    const __nv_bfloat16* A; long long rpad_a; int a_mn;
    const __nv_bfloat16* B; long long rpad_b; int b_mn;
    int M, N, K;
    int batch; long long bs_a, bs_b, bs_c, bs_bias, bs_mask;
    int splitk, k_chunk;
    int epi, accumulate;
    float* Cf; __nv_bfloat16* Cb; long long ldc;     // ldc: fp32 row stride, or rpad of an FB output
    int n_store;                                     // FB outputs: feature columns to write (>= N, zero filled)
    const float* bias;
    const __nv_bfloat16* mask; long long rpad_mask;
};

// NHWC-compact feature index n' = (y*35 + x)*32 + c  ->  reference column c*1225 + y*35 + x
__device__ __forceinline__ int nhwc_to_ref(int n) {
    const int c = n & 31, yx = n >> 5;
    return c * 1225 + yx;
}

template <int BN>
__global__ void __launch_bounds__(GT_THREADS, 1) gemm_tc_kernel(GemmTcArgs g) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int STAGES = gt_stages(BN);
    constexpr int A_BYTES = GT_BM * GT_BK * 2;
    constexpr int B_BYTES = BN * GT_BK * 2;
    constexpr int STAGE = A_BYTES + B_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* done = bars + 2 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * GT_BM, n0 = blockIdx.x * BN;
    const int z = blockIdx.z;
    const __nv_bfloat16* A = g.A;
    const __nv_bfloat16* B = g.B;
    int k_begin = 0, k_end = g.K;
    long long c_off = 0;
    if (g.splitk > 1) {
        k_begin = z * g.k_chunk;
        k_end = min(g.K, k_begin + g.k_chunk);
        c_off = (long long)z * g.bs_c;
    } else {
        A += (long long)z * g.bs_a;
        B += (long long)z * g.bs_b;
        c_off = (long long)z * g.bs_c;
    }
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nk = (k_end - k_begin + GT_BK - 1) / GT_BK;

    if (warp == 0) {
        // ------------------------------------------------ producer: bulk copies of FB pieces
        if (elect_one()) {
            // pieces per stage.  K-major: one piece per K unit (128 / BN rows x 16 B).
            // MN-major: one piece per M/N unit (BK rows x 16 B); units beyond the matrix are skipped
            // (their accumulator rows / columns are never stored).
            const int a_units = g.a_mn ? min(GT_BM / 8, (g.M - m0 + 7) / 8) : GT_BK / 8;
            const int b_units = g.b_mn ? min(BN / 8, (g.N - n0 + 7) / 8) : GT_BK / 8;
            const uint32_t a_piece = g.a_mn ? GT_BK * 16 : GT_BM * 16;
            const uint32_t b_piece = g.b_mn ? GT_BK * 16 : BN * 16;
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < nk; ++kb) {
                const int k0 = k_begin + kb * GT_BK;
                // K-major operands: only the K units that exist (K is padded to 16 in FB buffers)
                const int ku = min(GT_BK / 8, (k_end - k0 + 15) / 16 * 2);
                const int an = g.a_mn ? a_units : ku, bn_ = g.b_mn ? b_units : ku;
                mbar_wait(empty + stage, phase ^ 1);
                mbar_arrive_expect_tx(full + stage, an * a_piece + bn_ * b_piece);
                uint8_t* sa = smem + stage * STAGE;
                uint8_t* sb = sa + A_BYTES;
                for (int u = 0; u < an; ++u) {
                    const __nv_bfloat16* src = g.a_mn ? A + (((long long)(m0 / 8 + u)) * g.rpad_a + k0) * 8
                                                      : A + (((long long)(k0 / 8 + u)) * g.rpad_a + m0) * 8;
                    bulk_g2s(sa + u * a_piece, src, a_piece, full + stage);
                }
                for (int u = 0; u < bn_; ++u) {
                    const __nv_bfloat16* src = g.b_mn ? B + (((long long)(n0 / 8 + u)) * g.rpad_b + k0) * 8
                                                      : B + (((long long)(k0 / 8 + u)) * g.rpad_b + n0) * 8;
                    bulk_g2s(sb + u * b_piece, src, b_piece, full + stage);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ UMMA issuer
        const uint32_t idesc = make_idesc_bf16(GT_BM, BN, g.a_mn != 0, g.b_mn != 0);
        int stage = 0; uint32_t phase = 0;
        for (int kb = 0; kb < nk; ++kb) {
            mbar_wait(full + stage, phase);
            tc_fence_after();
            if (elect_one()) {
                const int k0 = k_begin + kb * GT_BK;
                const int ksteps = min(GT_BK / 16, (k_end - k0 + 15) / 16);
                const uint32_t a_addr = smem_u32(smem + stage * STAGE), b_addr = a_addr + A_BYTES;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint64_t da = g.a_mn ? make_smem_desc(a_addr + ks * 256, 128, GT_BK * 16)
                                               : make_smem_desc(a_addr + ks * 2 * GT_BM * 16, GT_BM * 16, 128);
                    const uint64_t db = g.b_mn ? make_smem_desc(b_addr + ks * 256, 128, GT_BK * 16)
                                               : make_smem_desc(b_addr + ks * 2 * BN * 16, BN * 16, 128);
                    umma_bf16(tmem_base, da, db, idesc, (kb | ks) ? 1u : 0u);
                }
                umma_commit(empty + stage);
                if (kb == nk - 1) umma_commit(done);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        // ------------------------------------------------ epilogue
        const int q = warp & 3;
        const int m = m0 + q * 32 + lane;
        const float* bias = g.bias ? g.bias + (g.splitk > 1 ? 0 : (long long)z * g.bs_bias) : nullptr;
        const __nv_bfloat16* mask = g.mask ? g.mask + (g.splitk > 1 ? 0 : (long long)z * g.bs_mask) : nullptr;
        mbar_wait(done, 0);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
            float v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
            const int nb = n0 + c0;
            if (m >= g.M) continue;
            if (g.epi == DRQ_TEPI_F32 || g.epi == DRQ_TEPI_TRUNK_WGRAD) {
                if (nb >= g.N) continue;
                const int nv = min(32, g.N - nb);
                float* crow = g.Cf + c_off + m * g.ldc;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j >= nv) continue;
                    float x = v[j];
                    if (g.epi == DRQ_TEPI_TRUNK_WGRAD) { crow[nhwc_to_ref(nb + j)] = x; continue; }
                    if (bias) x += __ldg(bias + nb + j);
                    if (g.accumulate) x += crow[nb + j];
                    crow[nb + j] = x;
                }
                continue;
            }
            if (g.epi == DRQ_TEPI_TRUNK_DGRAD) {
                if (nb >= g.N) continue;
                // columns nb..nb+31 = the 32 channels of feature pixel yx; mask by feature > 0 (FB feature
                // buffer: unit (nb/8 + c), row m) and scatter into conv4's WB gradient plane
                uint32_t packed[16];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 mv = __ldg(reinterpret_cast<const uint4*>(mask + (((long long)(nb / 8 + c)) * g.rpad_mask + m) * 8));
                    const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        packed[4 * c + j] = pack_bf16x2(bf16_lo(mw[j]) > 0.f ? v[8 * c + 2 * j] : 0.f,
                                                        bf16_hi(mw[j]) > 0.f ? v[8 * c + 2 * j + 1] : 0.f);
                }
                const int yx = nb >> 5;
                const int yy = yx / 35, xx = yx - yy * 35;
                const long long row = (long long)m * DRQ_PLB + DRQ_GUARD + yy * DRQ_PW + xx;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<uint4*>(g.Cb + (c * g.ldc + row) * 8) =
                        make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
                continue;
            }
            // FB bf16 output: unit (nb/8 + c), row m; columns >= N are written as zeros up to n_store
            if (nb >= g.n_store) continue;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int n8 = nb + 8 * c;
                if (n8 >= g.n_store) break;
                uint4 mv = make_uint4(0, 0, 0, 0);
                if (g.epi == DRQ_TEPI_MASK_BF16)
                    mv = __ldg(reinterpret_cast<const uint4*>(mask + (((long long)(n8 / 8)) * g.rpad_mask + m) * 8));
                const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
                uint32_t pk[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float lo = v[8 * c + 2 * j], hi = v[8 * c + 2 * j + 1];
                    const int n = n8 + 2 * j;
                    if (g.epi == DRQ_TEPI_RELU_BF16) {
                        lo = n < g.N ? fmaxf(lo + __ldg(bias + n), 0.f) : 0.f;
                        hi = n + 1 < g.N ? fmaxf(hi + __ldg(bias + n + 1), 0.f) : 0.f;
                    } else {
                        lo = (n < g.N && bf16_lo(mw[j]) > 0.f) ? lo : 0.f;
                        hi = (n + 1 < g.N && bf16_hi(mw[j]) > 0.f) ? hi : 0.f;
                    }
                    pk[j] = pack_bf16x2(lo, hi);
                }
                *reinterpret_cast<uint4*>(g.Cb + c_off + (((long long)(n8 / 8)) * g.ldc + m) * 8) =
                    make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
}

template <int BN>
static int launch_gemm_tc(const GemmTcArgs& g, cudaStream_t s) {
    constexpr size_t smem = gt_stages(BN) * (GT_BM * GT_BK * 2 + BN * GT_BK * 2) + (2 * gt_stages(BN) + 1) * 8 + 16;
    if (int rc = ensure_smem((const void*)gemm_tc_kernel<BN>, smem, "gemm_bf16")) return rc;
    dim3 grid((g.N + BN - 1) / BN, (g.M + GT_BM - 1) / GT_BM, g.splitk > 1 ? g.splitk : g.batch);
    gemm_tc_kernel<BN><<<grid, GT_THREADS, smem, s>>>(g);
    return check_launch("gemm_tc_kernel");
}

// fp32 nn.Linear weight [rows][cols] -> FB bf16 [ceil16(cols)/8][rpad][8] (zero padded)
__global__ void __launch_bounds__(256)
pack_linear_fb_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int rows, int cols, int rpad) {
    const int u = blockIdx.y;
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= rpad) return;
    uint32_t pk[4] = {0, 0, 0, 0};
    if (r < rows) {
        const float* src = w + (long long)r * cols + u * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = u * 8 + 2 * j;
            pk[j] = pack_bf16x2(c < cols ? src[2 * j] : 0.f, c + 1 < cols ? src[2 * j + 1] : 0.f);
        }
    }
    *reinterpret_cast<uint4*>(out + ((long long)u * rpad + r) * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// trunk weight fp32 [rows][32*1225] (reference NCHW-flatten columns c*1225+yx) -> FB bf16 with NHWC
// feature order n' = yx*32 + c: unit yx*4 + c/8.  32x32 shared-memory transpose per tile.
__global__ void __launch_bounds__(256)
pack_trunk_fb_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int rows, int rpad) {
    __shared__ float tile[32][33];
    const int r = blockIdx.y, yx0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* wr = w + (long long)r * DRQ_REPR_DIM;
    const bool live = r < rows;
#pragma unroll
    for (int c = ty; c < 32; c += 8) {
        const int yx = yx0 + tx;
        tile[c][tx] = (live && yx < 1225) ? wr[c * 1225 + yx] : 0.f;
    }
    __syncthreads();
    // thread -> (yx = yx0 + i, channel unit cu): 32 x 4 = 128 units per tile
    if (threadIdx.x < 128) {
        const int i = threadIdx.x >> 2, cu = threadIdx.x & 3;
        const int yx = yx0 + i;
        if (yx < 1225) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) pk[j] = pack_bf16x2(tile[cu * 8 + 2 * j][i], tile[cu * 8 + 2 * j + 1][i]);
            *reinterpret_cast<uint4*>(out + (((long long)(yx * 4 + cu)) * rpad + r) * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}

}  // namespace drq

using namespace drq;

extern "C" {

int drq_gemm_bf16(const uint16_t* A, int64_t rpad_a, int a_mn_major, const uint16_t* B, int64_t rpad_b,
                  int b_mn_major, void* C, int64_t ldc, int n_store, const float* bias, const uint16_t* mask,
                  int64_t rpad_mask, int M, int N, int K, int epilogue, int accumulate, int batch, int64_t bs_a,
                  int64_t bs_b, int64_t bs_c, int64_t bs_bias, int64_t bs_mask, int splitk, int bn, void* stream) {
    DRQ_REQUIRE(A && B && C, "gemm_bf16: null pointer");
    DRQ_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_bf16: bad dims");
    DRQ_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0, "gemm_bf16: operands must be 16-byte aligned");
    DRQ_REQUIRE(bs_a % 8 == 0 && bs_b % 8 == 0, "gemm_bf16: batch strides must keep 16-byte alignment");
    DRQ_REQUIRE(batch >= 1 && splitk >= 1 && !(batch > 1 && splitk > 1), "gemm_bf16: batch/splitk");
    DRQ_REQUIRE(epilogue >= DRQ_TEPI_F32 && epilogue <= DRQ_TEPI_TRUNK_DGRAD, "gemm_bf16: bad epilogue");
    DRQ_REQUIRE(!((epilogue == DRQ_TEPI_MASK_BF16 || epilogue == DRQ_TEPI_TRUNK_DGRAD) && !mask), "gemm_bf16: mask missing");
    DRQ_REQUIRE(!(epilogue == DRQ_TEPI_RELU_BF16 && !bias), "gemm_bf16: bias missing");
    DRQ_REQUIRE(!(splitk > 1 && epilogue != DRQ_TEPI_F32), "gemm_bf16: split-K writes fp32 partials");
    // row padding: K-major operands are copied 128 (A) / bn (B) rows at a time, MN-major ones 64 rows (K) at a time
    const int64_t need_a = a_mn_major ? (K + GT_BK - 1) / GT_BK * GT_BK : (M + GT_BM - 1) / GT_BM * GT_BM;
    const int64_t need_b = b_mn_major ? (K + GT_BK - 1) / GT_BK * GT_BK : (N + bn - 1) / bn * bn;
    DRQ_REQUIRE(rpad_a >= need_a && rpad_b >= need_b, "gemm_bf16: FB row padding too small (need %lld / %lld rows)",
                (long long)need_a, (long long)need_b);
    GemmTcArgs g{};
    g.A = reinterpret_cast<const __nv_bfloat16*>(A); g.rpad_a = rpad_a; g.a_mn = a_mn_major;
    g.B = reinterpret_cast<const __nv_bfloat16*>(B); g.rpad_b = rpad_b; g.b_mn = b_mn_major;
    g.M = M; g.N = N; g.K = K;
    g.batch = batch; g.bs_a = bs_a; g.bs_b = bs_b; g.bs_c = bs_c; g.bs_bias = bs_bias; g.bs_mask = bs_mask;
    g.splitk = splitk; g.k_chunk = K;
    if (splitk > 1) {
        int chunk = (K + splitk - 1) / splitk;
        chunk = (chunk + 127) / 128 * 128;
        g.k_chunk = chunk;
        DRQ_REQUIRE((long long)chunk * (splitk - 1) < K, "gemm_bf16: splitk %d leaves empty chunks for K=%d", splitk, K);
    }
    g.epi = epilogue; g.accumulate = accumulate;
    g.Cf = reinterpret_cast<float*>(C); g.Cb = reinterpret_cast<__nv_bfloat16*>(C); g.ldc = ldc;
    g.n_store = n_store > N ? n_store : N;
    g.bias = splitk > 1 ? nullptr : bias;
    g.mask = reinterpret_cast<const __nv_bfloat16*>(mask); g.rpad_mask = rpad_mask;
    cudaStream_t s = as_stream(stream);
    switch (bn) {
        case 32: return launch_gemm_tc<32>(g, s);
        case 64: return launch_gemm_tc<64>(g, s);
        case 128: return launch_gemm_tc<128>(g, s);
        default: set_error("gemm_bf16: bn must be 32, 64 or 128"); return DRQ_ERR_INVALID;
    }
}

int drq_pack_linear_fb(const float* w, uint16_t* out, int rows, int cols, int rpad, void* stream) {
    DRQ_REQUIRE(w && out && rows > 0 && cols > 0 && rpad >= rows, "pack_linear_fb: bad args");
    const int units = (cols + 15) / 16 * 2;
    pack_linear_fb_kernel<<<dim3((rpad + 255) / 256, units), 256, 0, as_stream(stream)>>>(
        w, reinterpret_cast<__nv_bfloat16*>(out), rows, cols, rpad);
    return check_launch("pack_linear_fb_kernel");
}

int drq_pack_trunk_fb(const float* w, uint16_t* out, int rows, int rpad, void* stream) {
    DRQ_REQUIRE(w && out && rows > 0 && rpad >= rows, "pack_trunk_fb: bad args");
    pack_trunk_fb_kernel<<<dim3((1225 + 31) / 32, rpad), 256, 0, as_stream(stream)>>>(
        w, reinterpret_cast<__nv_bfloat16*>(out), rows, rpad);
    return check_launch("pack_trunk_fb_kernel");
}

}  // extern "C"
