// Microbenchmark + correctness probe: tensor-map TMA (cp.async.bulk.tensor.5d) with traversal strides over the WB
// activation layout [4 channel blocks][N * 1776 + 128 pixel rows][8 ch] bf16.  A box of (8 ch, 42 x / 2, 14 y / 2, 4 cb)
// with element strides (1, 2, 2, 1) gathers one parity plane window [4 cb][7 block rows][21 blocks][16 B] - the A operand
// of a 2x2-output-block implicit GEMM (M128 N128/N64 UMMAs) - straight out of the unchanged row-major layout.
// Measures: are the gathered bytes right, and how many cycles does the TMA need per box (16-byte inner rows)?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../drqv2_b200/csrc/tc_common.cuh"
using namespace drq::tc;
namespace drq { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } int ensure_smem(const void*, size_t, const char*) { return 0; } }

constexpr int PLB = 1776, GUARD = 88, SLACK = 128, PW = 41;
constexpr int BX = 21, BY = 7;                       // blocks per box row / block rows per box
constexpr int BOX_BYTES = 4 * BY * BX * 16;          // 9408
constexpr int SLOT = (BOX_BYTES + 127) / 128 * 128;  // TMA destinations are 128-byte aligned

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar)) : "memory");
}

// every CTA loads `iters` rounds of the 4 parity windows of (image, tile) pairs; CTA 0 copies its last 4 boxes out
__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap map, int n_images, int iters, int boxes_in_flight,
                                                     uint16_t* out, long long* cycles) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar[8];
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(bar + i, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        uint32_t phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int issued = 0, waited = 0;
        const int total = iters * 4;
        while (waited < total) {
            while (issued < total && issued - waited < boxes_in_flight) {
                const int s = issued % 8, plane = issued & 3, it = issued >> 2;
                const int n = (blockIdx.x + it * gridDim.x) % n_images, ty = (it % 3) * 6;
                mbar_arrive_expect_tx(bar + s, BOX_BYTES);
                tma_load_5d(smem + s * SLOT, &map, 0, plane & 1, 2 * ty + (plane >> 1), 0, n, bar + s);
                ++issued;
            }
            const int s = waited % 8;
            mbar_wait(bar + s, phase[s]);
            phase[s] ^= 1;
            ++waited;
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    if (blockIdx.x == 0) {           // the last round's 4 boxes sit in slots (total-4 .. total-1) % 8
        const int total = iters * 4;
        for (int b = 0; b < 4; ++b) {
            const int s = (total - 4 + b) % 8;
            for (int i = threadIdx.x; i < BOX_BYTES / 2; i += 128) out[b * (BOX_BYTES / 2) + i] = reinterpret_cast<uint16_t*>(smem + s * SLOT)[i];
        }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int N = 512;
    const long long cs = (long long)N * PLB + SLACK;             // pixel rows per channel block
    const size_t elems = 4ull * cs * 8;
    std::vector<uint16_t> h(elems);
    for (size_t i = 0; i < elems; ++i) h[i] = (uint16_t)((i * 2654435761ull) >> 13);   // any bits: compared as bits
    uint16_t* d; cudaMalloc(&d, elems * 2); cudaMemcpy(d, h.data(), elems * 2, cudaMemcpyHostToDevice);
    EncodeFn encode = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q) != cudaSuccess || !encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap map;
    const cuuint64_t gdim[5] = {8, PW, PW, 4, (cuuint64_t)N};
    const cuuint64_t gstr[4] = {16, PW * 16, (cuuint64_t)cs * 16, (cuuint64_t)PLB * 16};   // bytes, dims 1..4
    const cuuint32_t box[5] = {8, 2 * BX, 2 * BY, 4, 1};
    const cuuint32_t estr[5] = {1, 2, 2, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d + GUARD * 8, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
    uint16_t* out; cudaMalloc(&out, 4 * BOX_BYTES); long long* cyc; cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * SLOT);
    for (int inflight : {1, 2, 4, 8}) {
        const int iters = 300;
        tma_kernel<<<148, 128, 8 * SLOT>>>(map, N, iters, inflight, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        long long hc[148]; cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
        const double per_box = (double)hc[0] / (iters * 4);
        printf("boxes in flight %d: %.0f cycles per box of %d B (%d rows of 16 B) -> %.1f B/cycle/SM, all 148 SMs loading\n", inflight, per_box,
               BOX_BYTES, BOX_BYTES / 16, BOX_BYTES / per_box);
        // check CTA 0's last round: it = iters-1, n = ((iters-1)*148) % N, ty = ((iters-1) % 3) * 6
        std::vector<uint16_t> got(4 * BOX_BYTES / 2);
        cudaMemcpy(got.data(), out, 4 * BOX_BYTES, cudaMemcpyDeviceToHost);
        const int it = iters - 1, n = (it * 148) % N, ty = (it % 3) * 6;
        long long bad = 0;
        for (int plane = 0; plane < 4; ++plane)
            for (int cb = 0; cb < 4; ++cb)
                for (int y = 0; y < BY; ++y)
                    for (int x = 0; x < BX; ++x)
                        for (int c = 0; c < 8; ++c) {
                            const int gy = 2 * (ty + y) + (plane >> 1), gx = 2 * x + (plane & 1);
                            uint16_t want = 0;                                  // out-of-bounds elements are zero-filled
                            if (gy < PW && gx < PW) want = h[((size_t)cb * cs + (size_t)n * PLB + GUARD + gy * PW + gx) * 8 + c];
                            const uint16_t g = got[(size_t)plane * (BOX_BYTES / 2) + ((cb * BY + y) * BX + x) * 8 + c];
                            bad += g != want;
                        }
        printf("  gathered planes of image %d, block rows %d..%d: %lld wrong elements of %d\n", n, ty, ty + BY - 1, bad, 4 * BOX_BYTES / 2);
    }
    return 0;
}
