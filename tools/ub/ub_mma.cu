// Microbenchmark: issue rate of tcgen05.mma (kind::f16, bf16, cta_group::1) from shared-memory operands
// as a function of N, operand major-ness and swizzle mode.  One CTA per SM, one issuing thread.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../drqv2_b200/csrc/tc_common.cuh"
using namespace drq::tc;
namespace drq { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } int ensure_smem(const void*, size_t, const char*) { return 0; } }

__device__ __forceinline__ uint64_t desc_sw(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | ((uint64_t)layout << 61);
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand read from tensor memory instead of shared memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// mode 0: K-major no swizzle ([unit][rows][16B]); 1: K-major 128B swizzle ([rows][128B], K=64 per row);
// 2: MN-major no swizzle for both; 3: A from TMEM (columns 128.. of the allocation), B K-major no swizzle
template <int mode, int distinct>
__global__ void __launch_bounds__(128, 2) mma_kernel(int M, int N, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + (i & 7);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(&slot, mode == 3 ? 512 : 256); tmem_relinquish(); }   // 2 CTAs per SM: 2 x 256 columns (mode 3: one CTA, 512)
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 48 * 1024;
        const uint32_t idesc = make_idesc_bf16(M, N, mode == 2, mode == 2);
        uint64_t da0, db0;
        if (mode == 0 || mode == 3) { da0 = desc_sw(a0, 128 * 16, 128, 0); db0 = desc_sw(b0, N * 16, 128, 0); }
        else if (mode == 1) { da0 = desc_sw(a0, 16, 1024, 2); db0 = desc_sw(b0, 16, 1024, 2); }
        else { da0 = desc_sw(a0, 128, 128 * 16, 0); db0 = desc_sw(b0, 128, 128 * 16, 0); }
        t0 = clock64();
        for (int i = 0; i < iters; i += 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t o = mode == 1 ? (j & 7) * 64 : (distinct == 9 ? ((j % 9) / 3) * 41 + (j % 9) % 3 : (j % 9) * 16);   // (address >> 4) offsets: conv taps / aligned
                if (mode == 3) umma_bf16_ts(tm, tm + 256 + (j & 7) * 8, db0 + (j & 1) * 32, idesc, 1u);   // 8 columns = one K16 slab of bf16 pairs
                else umma_bf16(tm, da0 + o, db0 + (mode == 1 ? 0 : (j & 1) * 32), idesc, 1u);
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, mode == 3 ? 512 : 256);
}

int main() {
    long long* out; cudaMalloc(&out, 148 * 8);
    printf("M N mode | cycles/MMA (grid 148)  -> MAC/clk/SM\n");
    for (int ctas : {1, 2}) {
    printf("---- %d CTA(s) per SM\n", ctas);
    for (int mode : {0, 3})
        for (int M : {64, 128})
            for (int N : {32, 64, 128, 256}) {
                if (mode == 2 && N == 256) continue;
                const int iters = 4000;
              for (int distinct : {9, 8}) {
                if (distinct == 8 && mode == 1) continue;
                if (mode == 3 && (ctas == 2 || distinct == 9 || N == 256)) continue;   // A in TMEM: one CTA per SM owns all 512 columns
                auto launch = [&](auto kern) {
                    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
                    kern<<<148 * ctas, 128, 100 * 1024>>>(M, N, iters, out);
                };
                if (mode == 0 && distinct == 9) launch(mma_kernel<0, 9>);
                else if (mode == 0) launch(mma_kernel<0, 8>);
                else launch(mma_kernel<3, 8>);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("M=%d N=%d mode=%d: %s\n", M, N, mode, cudaGetErrorString(e)); return 1; }
                long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                double c = (double)h[0] / iters;
                printf("%3d %3d %d %s | %7.1f  -> %.0f\n", M, N, mode, distinct == 9 ? "tap-offsets" : "aligned    ", c, (double)M * N * 16 / c);
              }
            }
    }
    return 0;
}
