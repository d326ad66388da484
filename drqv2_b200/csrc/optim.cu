// Multi-tensor Adam and the soft target update over the flat parameter arena.
// Reference: torch.optim.Adam x3 (drqv2.py:148-150) — single-tensor math of
// torch/optim/adam.py:457,476,531-547 — and utils.soft_update_params (utils.py:42-45).
#include "common.cuh"

namespace drq {

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float omb1, float b2,
                                          float omb2, float bc2_sqrt, float eps, float neg_step) {
    // exp_avg.lerp_(grad, 1-beta1): weight < 0.5 -> fma(weight, end - start, start)  (ATen Lerp.h)
    m = fmaf(omb1, __fsub_rn(g, m), m);
    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
    v = __fadd_rn(__fmul_rn(v, b2), __fmul_rn(__fmul_rn(omb2, g), g));
    // denom = (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps);
    // param.addcdiv_(exp_avg, denom, value=-step_size)
    p = __fadd_rn(p, __fmul_rn(neg_step, __fdiv_rn(m, denom)));
}

__device__ __forceinline__ void adam_range(float* __restrict__ p, const float* __restrict__ g,
                                           float* __restrict__ m, float* __restrict__ v, long long n,
                                           const float* __restrict__ sc, long long tid,
                                           long long nthreads) {
    const float omb1 = sc[0], b2 = sc[1], omb2 = sc[2], bc2s = sc[3], eps = sc[4], nstep = sc[5];
    const long long n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (long long i = tid; i < n4; i += nthreads) {
        float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
        adam_elem(pp.x, gg.x, mm.x, vv.x, omb1, b2, omb2, bc2s, eps, nstep);
        adam_elem(pp.y, gg.y, mm.y, vv.y, omb1, b2, omb2, bc2s, eps, nstep);
        adam_elem(pp.z, gg.z, mm.z, vv.z, omb1, b2, omb2, bc2s, eps, nstep);
        adam_elem(pp.w, gg.w, mm.w, vv.w, omb1, b2, omb2, bc2s, eps, nstep);
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
    }
    for (long long i = (n4 << 2) + tid; i < n; i += nthreads)
        adam_elem(p[i], g[i], m[i], v[i], omb1, b2, omb2, bc2s, eps, nstep);
}

__device__ __forceinline__ void ema_range(const float* __restrict__ src, float* __restrict__ dst,
                                          long long n, float tau, float omt, long long tid,
                                          long long nthreads) {
    const long long n4 = n >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (long long i = tid; i < n4; i += nthreads) {
        const float4 s = s4[i];
        float4 d = d4[i];
        d.x = __fadd_rn(__fmul_rn(tau, s.x), __fmul_rn(omt, d.x));   // utils.py:44-45
        d.y = __fadd_rn(__fmul_rn(tau, s.y), __fmul_rn(omt, d.y));
        d.z = __fadd_rn(__fmul_rn(tau, s.z), __fmul_rn(omt, d.z));
        d.w = __fadd_rn(__fmul_rn(tau, s.w), __fmul_rn(omt, d.w));
        d4[i] = d;
    }
    for (long long i = (n4 << 2) + tid; i < n; i += nthreads)
        dst[i] = __fadd_rn(__fmul_rn(tau, src[i]), __fmul_rn(omt, dst[i]));
}

// blocks [0, adam_blocks) run Adam, the rest run the EMA: one launch, two segments.
__global__ void __launch_bounds__(256)
adam_ema_kernel(float* p, const float* g, float* m, float* v, long long n_adam, const float* sc,
                const float* ema_src, float* ema_dst, long long n_ema, float tau, float omt,
                int adam_blocks) {
    pdl_trigger();
    pdl_wait();
    if ((int)blockIdx.x < adam_blocks) {
        adam_range(p, g, m, v, n_adam, sc, blockIdx.x * 256ll + threadIdx.x, adam_blocks * 256ll);
    } else {
        const int eb = gridDim.x - adam_blocks;
        ema_range(ema_src, ema_dst, n_ema, tau, omt, (blockIdx.x - adam_blocks) * 256ll + threadIdx.x,
                  eb * 256ll);
    }
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// Optional L2 residency of the Adam moments (drq_set_l2_persist): the optimiser kernels are launched with a
// persisting access-policy window over [base, base + bytes) so that m and v (8 of the 28 bytes per parameter
// read, 8 written) stay in the L2 set-aside between updates instead of streaming from HBM.
static const void* g_persist_base = nullptr;
static size_t g_persist_bytes = 0;

static int blocks_for(long long n) {
    long long b = (n / 4 + 255) / 256;
    if (b < 1) b = 1;
    if (b > 148 * 8) b = 148 * 8;   // 8 CTAs of 256 threads per SM on B200
    return (int)b;
}

}  // namespace drq

using namespace drq;

extern "C" {

int drq_adam_ema_step(float* p, const float* g, float* m, float* v, int64_t n_adam,
                      const float* scalars, const float* ema_src, float* ema_dst, int64_t n_ema,
                      float tau, float one_minus_tau, void* stream) {
    DRQ_REQUIRE(n_adam >= 0 && n_ema >= 0 && n_adam + n_ema > 0, "adam_ema: empty");
    DRQ_REQUIRE(n_adam == 0 || (p && g && m && v && scalars), "adam_ema: null adam pointer");
    DRQ_REQUIRE(n_ema == 0 || (ema_src && ema_dst), "adam_ema: null ema pointer");
    DRQ_REQUIRE(n_adam == 0 || (aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v)),
                "adam_ema: adam arenas must be 16-byte aligned");
    DRQ_REQUIRE(n_ema == 0 || (aligned16(ema_src) && aligned16(ema_dst)),
                "adam_ema: ema arenas must be 16-byte aligned");
    const int ab = n_adam ? blocks_for(n_adam) : 0;
    const int eb = n_ema ? blocks_for(n_ema) : 0;
    if (g_persist_bytes) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(ab + eb); cfg.blockDim = dim3(256); cfg.stream = as_stream(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = const_cast<void*>(g_persist_base);
        attr[0].val.accessPolicyWindow.num_bytes = g_persist_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, adam_ema_kernel, p, (const float*)g, m, v, (long long)n_adam, scalars, ema_src, ema_dst,
                           (long long)n_ema, tau, one_minus_tau, ab);
    } else {
        launch_k(adam_ema_kernel, ab + eb, 256, 0, as_stream(stream), p, g, m, v, n_adam, scalars, ema_src, ema_dst,
                 n_ema, tau, one_minus_tau, ab);
    }
    return check_launch("adam_ema_kernel");
}

int drq_set_l2_persist(const void* base, int64_t bytes) {
    if (!base || bytes <= 0) { g_persist_base = nullptr; g_persist_bytes = 0; return DRQ_OK; }
    int dev = 0, max_persist = 0, max_window = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    size_t want = (size_t)bytes;
    if (want > (size_t)max_persist) want = (size_t)max_persist;
    if (want > (size_t)max_window) want = (size_t)max_window;
    if (want == 0) { set_error("l2_persist: device has no persisting L2"); return DRQ_ERR_CUDA; }
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) {
        cudaGetLastError();
        set_error("l2_persist: cudaDeviceSetLimit failed");
        return DRQ_ERR_CUDA;
    }
    g_persist_base = base; g_persist_bytes = want;
    return DRQ_OK;
}

int drq_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const float* scalars,
                  void* stream) {
    return drq_adam_ema_step(p, g, m, v, n, scalars, nullptr, nullptr, 0, 0.f, 0.f, stream);
}

int drq_soft_update(const float* p, float* tp, int64_t n, float tau, float one_minus_tau, void* stream) {
    return drq_adam_ema_step(nullptr, nullptr, nullptr, nullptr, 0, nullptr, p, tp, n, tau,
                             one_minus_tau, stream);
}

}  // extern "C"
