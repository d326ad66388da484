// Microbenchmark: tcgen05.mma.cta_group::2 (one instruction drives the tensor cores of both SMs of a CTA pair:
// M = 256 = 128 rows from each CTA's shared memory, N split over the pair) against cta_group::1 at the conv shapes.
// Question: is the ~45-cycle floor of an M128 N32 K16 UMMA per instruction (then a pair
// instruction doubles the work per issue slot) or per SM?
#include <cstdio>
#include <cuda_runtime.h>
#include "../../drqv2_b200/csrc/tc_common.cuh"
using namespace drq::tc;
namespace drq { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } int ensure_smem(const void*, size_t, const char*) { return 0; } }

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"(mask) : "memory");
}

// PAIR = 1: clusters of 2, the leader issues M256 x N UMMAs; PAIR = 0: every CTA issues M128 x N UMMAs (reference)
template <int PAIR>
__global__ void __launch_bounds__(128, 1) mma2_kernel(int N, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u + (i & 7);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) {
        if (PAIR) { tmem_alloc2(&slot, 256); tmem_relinquish2(); }
        else { tmem_alloc(&slot, 256); tmem_relinquish(); }
    }
    fence_proxy_async();
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const bool leader = !PAIR || cluster_rank() == 0;
    if (threadIdx.x == 0 && leader) {
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 48 * 1024;
        // K-major, no swizzle: A [unit][128 rows][16 B] in each CTA; B [unit][N (pair: N/2 per CTA) rows][16 B]
        const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, N, false, false);
        const int nb = PAIR ? N / 2 : N;
        const uint64_t da0 = make_smem_desc(a0, 128 * 16, 128), db0 = make_smem_desc(b0, nb * 16, 128);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t o = ((j % 9) / 3) * 41 + (j % 9) % 3;      // the conv kernels' tap offsets (address >> 4)
                if (PAIR) umma2_bf16(tm, da0 + o, db0 + (j & 1) * 32, idesc, 1u);
                else umma_bf16(tm, da0 + o, db0 + (j & 1) * 32, idesc, 1u);
            }
        }
        if (PAIR) umma2_commit(&bar, 1); else umma_commit(&bar);
        mbar_wait(&bar, 0);
        out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (threadIdx.x < 32) { if (PAIR) tmem_dealloc2(tm, 256); else tmem_dealloc(tm, 256); }
}

template <int PAIR>
static bool run(int N, int iters, long long* out) {
    cudaFuncSetAttribute(mma2_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 100 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, mma2_kernel<PAIR>, N, iters, out);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("pair=%d N=%d: %s\n", PAIR, N, cudaGetErrorString(e)); return false; }
    long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    const double c = (double)h[0] / iters;
    const double macs = (double)(PAIR ? 256 : 128) * N * 16;
    printf("cta_group::%d  M=%3d N=%3d | %7.1f cycles per UMMA -> %.0f MAC/clk per SM\n", PAIR ? 2 : 1, PAIR ? 256 : 128, N, c,
           macs / c / (PAIR ? 2 : 1));
    return true;
}

int main() {
    long long* out; cudaMalloc(&out, 148 * 8); cudaMemset(out, 0, 148 * 8);
    const int iters = 4000;
    for (int N : {32, 64, 128, 256}) if (!run<0>(N, iters, out)) return 1;
    for (int N : {32, 64, 128, 256}) if (!run<1>(N, iters, out)) return 1;
    return 0;
}
