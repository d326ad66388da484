// Microbenchmark: per-SM throughput of 1-D cp.async.bulk (global -> smem) as a function of piece
// size and pipeline depth.  One elected thread issues; one consumer thread waits and releases.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../drqv2_b200/csrc/tc_common.cuh"
using namespace drq::tc;

namespace drq { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } int ensure_smem(const void*, size_t, const char*) { return 0; } }

__global__ void __launch_bounds__(64, 1) bulk_kernel(const uint8_t* src, size_t src_bytes, int piece, int pieces_per_stage,
                                                     int stages, int iters, int issuers, unsigned long long* cycles) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int stage_bytes = piece * pieces_per_stage;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
    uint64_t* empty = full + stages;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        fence_barrier_init();
    }
    __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x < 32) {
        // producer warp: lane l issues pieces l, l+issuers, ...
        int stage = 0; uint32_t phase = 0;
        size_t off = ((size_t)blockIdx.x * 7919 * 4096) % (src_bytes - (size_t)stage_bytes * 2);
        for (int it = 0; it < iters; ++it) {
            mbar_wait(empty + stage, phase ^ 1);
            if (threadIdx.x == 0) mbar_arrive_expect_tx(full + stage, stage_bytes);
            __syncwarp();
            for (int p = threadIdx.x; p < pieces_per_stage; p += issuers)
                if (threadIdx.x < issuers)
                    bulk_g2s(smem + stage * stage_bytes + p * piece, src + off + (size_t)p * piece, piece, full + stage);
            off += stage_bytes;
            if (off + stage_bytes > src_bytes) off = 0;
            if (++stage == stages) { stage = 0; phase ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(full + stage, phase);
            mbar_arrive(empty + stage);
            if (++stage == stages) { stage = 0; phase ^= 1; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
    size_t src_bytes = 96ull << 20;
    uint8_t* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
    unsigned long long* cyc; cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int grids[] = {148};
    int pieces[] = {512, 1024, 2048, 3392, 4096, 8192};
    printf("grid piece pps stages issuers | us  GB/s/SM  ns/copy  total GB/s\n");
    for (int g : grids) for (int piece : pieces) for (int issuers : {4, 16, 32}) {
        const int stage_target = 24576;
        int pps = stage_target / piece; if (pps < 1) pps = 1;
        int stages = 8; while ((size_t)stages * pps * piece > 190 * 1024) --stages;
        int iters = 256;
        size_t smem = (size_t)stages * pps * piece + stages * 16 + 64;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        bulk_kernel<<<g, 64, smem>>>(src, src_bytes, piece, pps, stages, iters, issuers, cyc);
        cudaEventRecord(e0);
        bulk_kernel<<<g, 64, smem>>>(src, src_bytes, piece, pps, stages, iters, issuers, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
        double bytes = (double)iters * pps * piece;
        printf("%4d %6d %3d %2d %2d | %8.1f %8.1f %8.1f %9.1f\n", g, piece, pps, stages, issuers, ms * 1e3, bytes / (ms * 1e-3) / 1e9,
               ms * 1e6 / (iters * pps), bytes * g / (ms * 1e-3) / 1e9);
    }
    return 0;
}
