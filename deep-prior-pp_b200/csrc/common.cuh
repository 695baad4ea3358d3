// Shared helpers for the dpp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/dpp_b200.h"

namespace dpp {

extern thread_local char g_err[512];

inline int fail(int code, const char *fmt, const char *a = "", const char *b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

#define DPP_CHECK_ARG(cond)                                                                  \
    do {                                                                                     \
        if (!(cond)) return dpp::fail(DPP_EINVAL, "%s: argument check failed: %s", __func__, #cond); \
    } while (0)

#define DPP_CUDA(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) return dpp::fail(DPP_ECUDA, "%s: %s", __func__, cudaGetErrorString(e__)); \
    } while (0)

#define DPP_LAUNCH_CHECK() DPP_CUDA(cudaGetLastError())

inline cudaStream_t S(void *s) { return reinterpret_cast<cudaStream_t>(s); }

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch ------------------------------------------------------
// The step is a chain of ~285 short kernels.  A kernel launched through launch_pdl() may start while its
// stream predecessor is still draining: it runs its private set-up (barriers, TMEM allocation, tables) and then
// blocks in pdl_wait() until the predecessor grid has completed and its memory is visible.  Contract: a kernel
// launched with launch_pdl() touches NO global memory before pdl_wait().  pdl_trigger() lets the next kernel
// start its own set-up.  Both are no-ops in a kernel launched the ordinary way.  DPP_PDL (bit 0: main chain,
// bit 1: backward-weights kernels) switches the launch attribute; CUDA-graph capture keeps the programmatic edges.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifndef DPP_PDL_DEFAULT
#define DPP_PDL_DEFAULT 3
#endif
int pdl_mode();     // misc.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int bit, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args &&...args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl_mode() & bit) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- BN prologue coefficients ---------------------------------------------------------
// a = max(x*scale + shift, 0) with scale = gamma*inv_std, shift = beta - mean*scale.
// mean / inv_std come from fp64 batch sums (train) or stored running stats (test).
__device__ __forceinline__ void bn_mean_istd(const dpp_bn_ref &bn, int c, int C, float &mean, float &inv_std) {
    if (bn.sums != nullptr) {
        double m = bn.sums[c] / bn.count;
        double var = bn.sums[C + c] / bn.count - m * m;
        if (var < 0.0) var = 0.0;
        mean = (float)m;
        inv_std = (float)(1.0 / sqrt(var + (double)bn.eps));
    } else {
        mean = bn.mean[c];
        inv_std = bn.inv_std[c];
    }
}

__device__ __forceinline__ void bn_scale_shift(const dpp_bn_ref &bn, int c, int C, float &scale, float &shift) {
    float mean, istd;
    bn_mean_istd(bn, c, C, mean, istd);
    scale = bn.gamma[c] * istd;
    shift = bn.beta[c] - mean * scale;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace dpp
