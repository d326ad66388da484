// bf16 tensor-core GEMM (tcgen05 + TMEM) for the heads in bf16 mode: trunk Linear(39200->F)
// forward (split-K) / weight gradient / data gradient, and the actor / twin-Q MLP layers.
// Reference ops: nn.Linear forward/backward of drqv2.py:74-81,100-111.
//
//   C[M,N] = sum_k A(m,k) * B(n,k),  bf16 operands, fp32 accumulation in TMEM.
//
// Each operand is either K-major (memory [row][k], k contiguous) or MN-major (memory [k][row],
// row contiguous); leading dimensions are multiples of 8 elements so every 16-byte unit is
// aligned.  128 threads: all of them stage 128xBK / BNxBK operand tiles into the no-swizzle
// canonical UMMA layout ([16-byte unit][line][16 B], which is the same address formula for
// both major-nesses), thread 0 issues the UMMAs, then the 4 warps read their TMEM lane
// quarters for the fused epilogue.
#include "tc_common.cuh"

namespace drq {

using namespace tc;

constexpr int GT_BM = 128, GT_BK = 128, GT_STAGES = 3, GT_LOADERS = 128, GT_THREADS = 160;

struct GemmTcArgs {
    const __nv_bfloat16* A; long long lda; int a_mn;
    const __nv_bfloat16* B; long long ldb; int b_mn;
    int M, N, K;
    int batch; long long bs_a, bs_b, bs_c, bs_bias, bs_mask;
    int splitk, k_chunk;
    int epi, accumulate;
    float* Cf; __nv_bfloat16* Cb; long long ldc;
    const float* bias;
    const __nv_bfloat16* mask; long long ldmask;
};

// stage one operand tile: `nlines` lines x `nunits` 16-byte units.
//   K-major : line = row (valid < rows_valid), unit = k/8   (valid while k < k_valid)
//   MN-major: line = k   (valid < k_valid),    unit = row/8 (valid while row < rows_valid)
// global unit address = base + line*ld + unit*8 ; smem = unit*(nlines*16) + line*16
__device__ __forceinline__ void cp_async16(uint8_t* smem_dst, const void* gsrc, bool valid) {
    const uint32_t n = valid ? 16u : 0u;      // src-size 0 => 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(n)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NLINES, int NUNITS>
__device__ __forceinline__ void stage_tile(uint8_t* smem_tile, const __nv_bfloat16* base, long long ld,
                                           int lines_valid, int units_valid_elems, int tid) {
    constexpr int TOTAL = NLINES * NUNITS;
#pragma unroll
    for (int i = 0; i < TOTAL / GT_LOADERS; ++i) {
        const int idx = tid + i * GT_LOADERS;
        const int line_lo = idx & 7;
        const int u = (idx >> 3) % NUNITS;
        const int line = ((idx >> 3) / NUNITS) * 8 + line_lo;
        const bool ok = line < lines_valid && u * 8 < units_valid_elems;
        cp_async16(smem_tile + u * (NLINES * 16) + line * 16, ok ? (const void*)(base + line * ld + u * 8) : (const void*)base, ok);
    }
}

// NHWC-compact feature index n' = (y*35 + x)*32 + c
__device__ __forceinline__ int nhwc_to_ref(int n) {          // -> c*1225 + y*35 + x
    const int c = n & 31, yx = n >> 5;
    return c * 1225 + yx;
}

template <int BN>
__global__ void __launch_bounds__(GT_THREADS) gemm_tc_kernel(GemmTcArgs g) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int A_BYTES = GT_BM * GT_BK * 2;
    constexpr int B_BYTES = BN * GT_BK * 2;
    constexpr int STAGE = A_BYTES + B_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GT_STAGES * STAGE);
    uint64_t* full = bars;                 // loaders -> MMA   (count 128)
    uint64_t* empty = bars + GT_STAGES;    // MMA commit -> loaders
    uint64_t* done = bars + 2 * GT_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GT_STAGES + 1);

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
        for (int i = 0; i < GT_STAGES; ++i) { mbar_init(full + i, GT_LOADERS); mbar_init(empty + i, 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 4) {
        tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = make_idesc_bf16(GT_BM, BN, g.a_mn != 0, g.b_mn != 0);
    const int nk = (k_end - k_begin + GT_BK - 1) / GT_BK;

    if (warp < 4) {
        // ------------------------------------------------ loaders: cp.async straight into the UMMA layout,
        // GT_STAGES-1 stages of loads in flight per thread
        auto issue = [&](int kb) {
            const int st = kb % GT_STAGES;
            const int k0 = k_begin + kb * GT_BK;
            const int kv = k_end - k0;
            uint8_t* sa = smem + st * STAGE;
            uint8_t* sb = sa + A_BYTES;
            if (!g.a_mn) stage_tile<GT_BM, GT_BK / 8>(sa, A + (long long)m0 * g.lda + k0, g.lda, g.M - m0, kv, tid);
            else         stage_tile<GT_BK, GT_BM / 8>(sa, A + (long long)k0 * g.lda + m0, g.lda, kv, g.M - m0, tid);
            if (!g.b_mn) stage_tile<BN, GT_BK / 8>(sb, B + (long long)n0 * g.ldb + k0, g.ldb, g.N - n0, kv, tid);
            else         stage_tile<GT_BK, BN / 8>(sb, B + (long long)k0 * g.ldb + n0, g.ldb, kv, g.N - n0, tid);
        };
        for (int kb = 0; kb < GT_STAGES - 1; ++kb) {
            if (kb < nk) issue(kb);
            cp_async_commit();
        }
        for (int kb = 0; kb < nk; ++kb) {
            const int nxt = kb + GT_STAGES - 1;
            if (nxt < nk) {
                mbar_wait(empty + nxt % GT_STAGES, ((nxt / GT_STAGES) & 1) ^ 1);
                issue(nxt);
            }
            cp_async_commit();
            cp_async_wait<GT_STAGES - 1>();      // the group of k-block kb has landed
            fence_proxy_async();                 // generic-proxy writes -> visible to the UMMA (async proxy)
            mbar_arrive(full + kb % GT_STAGES);
        }
    } else if (lane == 0) {
        // ------------------------------------------------ UMMA issuer
        for (int kb = 0; kb < nk; ++kb) {
            const int st = kb % GT_STAGES;
            mbar_wait(full + st, (kb / GT_STAGES) & 1);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + st * STAGE), b_addr = a_addr + A_BYTES;
#pragma unroll
            for (int ks = 0; ks < GT_BK / 16; ++ks) {
                const uint64_t da = g.a_mn ? make_smem_desc(a_addr + ks * 256, 128, GT_BK * 16)
                                           : make_smem_desc(a_addr + ks * 2 * GT_BM * 16, GT_BM * 16, 128);
                const uint64_t db = g.b_mn ? make_smem_desc(b_addr + ks * 256, 128, GT_BK * 16)
                                           : make_smem_desc(b_addr + ks * 2 * BN * 16, BN * 16, 128);
                umma_bf16(tmem_base, da, db, idesc, (kb | ks) ? 1u : 0u);
            }
            umma_commit(empty + st);
            if (kb == nk - 1) umma_commit(done);
        }
    }
    if (warp >= 4) {
        // the issuer warp takes no part in the epilogue; it only frees TMEM at the end
        tc_fence_before();
        __syncthreads();
        if (warp == 4) tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
        return;
    }
    mbar_wait(done, 0);
    tc_fence_after();

    // ---------------------------------------------------------------- epilogue
    const int m = m0 + warp * 32 + lane;
    const float* bias = g.bias ? g.bias + (g.splitk > 1 ? 0 : (long long)z * g.bs_bias) : nullptr;
    const __nv_bfloat16* mask = g.mask ? g.mask + (g.splitk > 1 ? 0 : (long long)z * g.bs_mask) : nullptr;
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        const int nb = n0 + c0;
        if (m >= g.M || nb >= g.N) continue;
        const int nv = min(32, g.N - nb);              // valid columns of this 32-wide chunk
        if (g.epi == DRQ_TEPI_F32 || g.epi == DRQ_TEPI_TRUNK_WGRAD) {
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
        // bf16 outputs: apply bias+ReLU or the ReLU mask, pack, store 16 bytes at a time
        if (g.epi == DRQ_TEPI_RELU_BF16) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (j < nv) ? fmaxf(v[j] + __ldg(bias + nb + j), 0.f) : 0.f;
        } else {
            const __nv_bfloat16* mrow = mask + m * g.ldmask + nb;
            if (nv == 32) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 mv = __ldg(reinterpret_cast<const uint4*>(mrow + 8 * c));
                    const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (!(bf16_lo(mw[j]) > 0.f)) v[8 * c + 2 * j] = 0.f;
                        if (!(bf16_hi(mw[j]) > 0.f)) v[8 * c + 2 * j + 1] = 0.f;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    v[j] = (j < nv && __bfloat162float(mrow[j]) > 0.f) ? v[j] : 0.f;
            }
        }
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) packed[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
        if (g.epi == DRQ_TEPI_TRUNK_DGRAD) {
            // columns nb..nb+31 = the 32 channels of feature pixel yx: scatter into conv4's WB gradient plane
            const int yx = nb >> 5;
            const int yy = yx / 35, xx = yx - yy * 35;
            const long long row = (long long)m * DRQ_PLB + DRQ_GUARD + yy * DRQ_PW + xx;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(g.Cb + (c * g.ldc + row) * 8) =
                    make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
            continue;
        }
        __nv_bfloat16* crow = g.Cb + c_off + m * g.ldc + nb;
        if (nv == 32) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(crow + 8 * c) =
                    make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < nv) crow[j] = __float2bfloat16_rn(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
}

template <int BN>
static int launch_gemm_tc(const GemmTcArgs& g, cudaStream_t s) {
    constexpr size_t smem = GT_STAGES * (GT_BM * GT_BK * 2 + BN * GT_BK * 2) + (2 * GT_STAGES + 1) * 8 + 16;
    if (int rc = ensure_smem((const void*)gemm_tc_kernel<BN>, smem, "gemm_bf16")) return rc;
    dim3 grid((g.N + BN - 1) / BN, (g.M + GT_BM - 1) / GT_BM, g.splitk > 1 ? g.splitk : g.batch);
    gemm_tc_kernel<BN><<<grid, GT_THREADS, smem, s>>>(g);
    return check_launch("gemm_tc_kernel");
}

// fp32 [rows][cols] -> bf16 [rows][ld] (zero padded); nhwc_permute: column k of the output is
// the NHWC feature index, read from the reference's NCHW-flatten column.
__global__ void pack_linear_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int rows, int cols,
                                   int ld, int nhwc_permute) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)rows * ld) return;
    const int r = (int)(i / ld), c = (int)(i - (long long)r * ld);
    float v = 0.f;
    if (c < cols) v = w[(long long)r * cols + (nhwc_permute ? nhwc_to_ref(c) : c)];
    out[i] = __float2bfloat16_rn(v);
}

}  // namespace drq

using namespace drq;

extern "C" {

int drq_gemm_bf16(const uint16_t* A, int64_t lda, int a_mn_major, const uint16_t* B, int64_t ldb, int b_mn_major,
                  void* C, int64_t ldc, const float* bias, const uint16_t* mask, int64_t ldmask, int M, int N, int K,
                  int epilogue, int accumulate, int batch, int64_t bs_a, int64_t bs_b, int64_t bs_c, int64_t bs_bias,
                  int64_t bs_mask, int splitk, int bn, void* stream) {
    DRQ_REQUIRE(A && B && C, "gemm_bf16: null pointer");
    DRQ_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_bf16: bad dims");
    DRQ_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0,
                "gemm_bf16: operands need 16-byte aligned rows (ld %% 8 == 0)");
    DRQ_REQUIRE(bs_a % 8 == 0 && bs_b % 8 == 0, "gemm_bf16: batch strides must keep 16-byte alignment");
    DRQ_REQUIRE(batch >= 1 && splitk >= 1 && !(batch > 1 && splitk > 1), "gemm_bf16: batch/splitk");
    DRQ_REQUIRE(epilogue >= DRQ_TEPI_F32 && epilogue <= DRQ_TEPI_TRUNK_DGRAD, "gemm_bf16: bad epilogue");
    DRQ_REQUIRE(!((epilogue == DRQ_TEPI_MASK_BF16 || epilogue == DRQ_TEPI_TRUNK_DGRAD) && !mask), "gemm_bf16: mask missing");
    DRQ_REQUIRE(!(epilogue == DRQ_TEPI_RELU_BF16 && !bias), "gemm_bf16: bias missing");
    DRQ_REQUIRE(!(splitk > 1 && epilogue != DRQ_TEPI_F32), "gemm_bf16: split-K writes fp32 partials");
    GemmTcArgs g{};
    g.A = reinterpret_cast<const __nv_bfloat16*>(A); g.lda = lda; g.a_mn = a_mn_major;
    g.B = reinterpret_cast<const __nv_bfloat16*>(B); g.ldb = ldb; g.b_mn = b_mn_major;
    g.M = M; g.N = N; g.K = K;
    g.batch = batch; g.bs_a = bs_a; g.bs_b = bs_b; g.bs_c = bs_c; g.bs_bias = bs_bias; g.bs_mask = bs_mask;
    g.splitk = splitk; g.k_chunk = K;
    if (splitk > 1) {
        int chunk = (K + splitk - 1) / splitk;
        chunk = (chunk + GT_BK - 1) / GT_BK * GT_BK;
        g.k_chunk = chunk;
        DRQ_REQUIRE((long long)chunk * (splitk - 1) < K, "gemm_bf16: splitk %d leaves empty chunks for K=%d", splitk, K);
        g.bias = nullptr;
    }
    g.epi = epilogue; g.accumulate = accumulate;
    g.Cf = reinterpret_cast<float*>(C); g.Cb = reinterpret_cast<__nv_bfloat16*>(C); g.ldc = ldc;
    g.bias = splitk > 1 ? nullptr : bias;
    g.mask = reinterpret_cast<const __nv_bfloat16*>(mask); g.ldmask = ldmask;
    cudaStream_t s = as_stream(stream);
    switch (bn) {
        case 32: return launch_gemm_tc<32>(g, s);
        case 64: return launch_gemm_tc<64>(g, s);
        case 128: return launch_gemm_tc<128>(g, s);
        default: set_error("gemm_bf16: bn must be 32, 64 or 128"); return DRQ_ERR_INVALID;
    }
}

int drq_pack_linear_bf16(const float* w, uint16_t* out, int rows, int cols, int ld, int nhwc_permute, void* stream) {
    DRQ_REQUIRE(w && out && rows > 0 && cols > 0 && ld >= cols && ld % 8 == 0, "pack_linear: bad args");
    DRQ_REQUIRE(!(nhwc_permute && cols != DRQ_REPR_DIM), "pack_linear: permute needs cols = 39200");
    const long long n = (long long)rows * ld;
    pack_linear_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
        w, reinterpret_cast<__nv_bfloat16*>(out), rows, cols, ld, nhwc_permute);
    return check_launch("pack_linear_kernel");
}

}  // extern "C"
