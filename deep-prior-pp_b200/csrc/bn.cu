// BatchNorm pieces that are not fused into a conv: backward apply (full backward through the
// batch statistics, as Theano's T.grad produces for net/batchnormlayer.py:154-192), the
// materialised BN+ReLU in front of the FC stack (net/resnet.py:138-141), its backward reduce, and
// the running-statistics EMA (net/batchnormlayer.py:164-172, EMA of mean and of INV_STD).
// All are HBM-bound elementwise passes over [pixels, C] with C contiguous.
#include "common.cuh"

using namespace dpp;

namespace {

constexpr int BT = 256;

// thread -> (channel quad cq = tid % (C/4), first pixel tid / (C/4)); C/4 divides 256.
__global__ void __launch_bounds__(BT)
k_bn_bwd_apply(const float *__restrict__ dz, const float *__restrict__ x, dpp_bn_ref bn,
               const double *__restrict__ dz_stats, const float *__restrict__ skip, float *__restrict__ dx,
               float *__restrict__ dgamma, float *__restrict__ dbeta, double *__restrict__ dbias_stats,
               int64_t pixels, int C, float pscale) {
    __shared__ float s_k1[256], s_mdz[256], s_mdzx[256], s_mean[256], s_istd[256];
    __shared__ float s_red[BT][4];
    const int tid = threadIdx.x;
    pdl_trigger();
    pdl_wait();
    for (int c = tid; c < C; c += BT) {
        float mean, istd;
        bn_mean_istd(bn, c, C, mean, istd);
        s_mean[c] = mean; s_istd[c] = istd;
        s_k1[c] = bn.gamma[c] * istd;
        s_mdz[c] = (float)(dz_stats[c] / bn.count);
        s_mdzx[c] = (float)(dz_stats[C + c] / bn.count);
        if (blockIdx.x == 0) {
            if (dbeta) dbeta[c] += pscale * (float)dz_stats[c];
            if (dgamma) dgamma[c] += pscale * (float)dz_stats[C + c];
        }
    }
    __syncthreads();
    const int cq4 = C / 4;
    const int cq = tid % cq4, c = cq * 4;
    const int prow = tid / cq4, pstep = BT / cq4;
    float bs[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t p = (int64_t)blockIdx.x * pstep + prow; p < pixels; p += (int64_t)gridDim.x * pstep) {
        size_t o = (size_t)p * C + c;
        float4 g = *reinterpret_cast<const float4 *>(dz + o);
        float4 xv = *reinterpret_cast<const float4 *>(x + o);
        float gr[4] = {g.x, g.y, g.z, g.w}, xr[4] = {xv.x, xv.y, xv.z, xv.w}, r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float xh = (xr[j] - s_mean[c + j]) * s_istd[c + j];
            r[j] = s_k1[c + j] * (gr[j] - s_mdz[c + j] - xh * s_mdzx[c + j]);
        }
        if (skip) {
            float4 sk = *reinterpret_cast<const float4 *>(skip + o);
            r[0] += sk.x; r[1] += sk.y; r[2] += sk.z; r[3] += sk.w;
        }
        *reinterpret_cast<float4 *>(dx + o) = make_float4(r[0], r[1], r[2], r[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) bs[j] += r[j];
    }
    if (dbias_stats) {
#pragma unroll
        for (int j = 0; j < 4; ++j) s_red[tid][j] = bs[j];
        __syncthreads();
        if (tid < cq4) {
            double t[4] = {0, 0, 0, 0};
            for (int i = tid; i < BT; i += cq4)
#pragma unroll
                for (int j = 0; j < 4; ++j) t[j] += (double)s_red[i][j];
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(&dbias_stats[tid * 4 + j], t[j]);
        }
    }
}

__global__ void __launch_bounds__(BT)
k_bn_apply(const float *__restrict__ x, dpp_bn_ref bn, float *__restrict__ y, int64_t pixels, int C) {
    __shared__ float s_scale[256], s_shift[256];
    for (int c = threadIdx.x; c < C; c += BT) bn_scale_shift(bn, c, C, s_scale[c], s_shift[c]);
    __syncthreads();
    const int cq4 = C / 4;
    const int c = (threadIdx.x % cq4) * 4;
    const int prow = threadIdx.x / cq4, pstep = BT / cq4;
    for (int64_t p = (int64_t)blockIdx.x * pstep + prow; p < pixels; p += (int64_t)gridDim.x * pstep) {
        size_t o = (size_t)p * C + c;
        float4 v = *reinterpret_cast<const float4 *>(x + o);
        v.x = fmaf(v.x, s_scale[c], s_shift[c]);
        v.y = fmaf(v.y, s_scale[c + 1], s_shift[c + 1]);
        v.z = fmaf(v.z, s_scale[c + 2], s_shift[c + 2]);
        v.w = fmaf(v.w, s_scale[c + 3], s_shift[c + 3]);
        if (bn.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        *reinterpret_cast<float4 *>(y + o) = v;
    }
}

__global__ void __launch_bounds__(BT)
k_bn_relu_bwd_reduce(const float *__restrict__ dy, const float *__restrict__ x, dpp_bn_ref bn,
                     float *__restrict__ dz, double *__restrict__ dz_stats, int64_t pixels, int C) {
    __shared__ float s_scale[256], s_shift[256], s_mean[256], s_istd[256];
    __shared__ float s_red[2][BT][4];
    const int tid = threadIdx.x;
    for (int c = tid; c < C; c += BT) {
        float mean, istd;
        bn_mean_istd(bn, c, C, mean, istd);
        s_mean[c] = mean; s_istd[c] = istd;
        float sc = bn.gamma[c] * istd;
        s_scale[c] = sc; s_shift[c] = bn.beta[c] - mean * sc;
    }
    __syncthreads();
    const int cq4 = C / 4;
    const int c = (tid % cq4) * 4;
    const int prow = tid / cq4, pstep = BT / cq4;
    float a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
    for (int64_t p = (int64_t)blockIdx.x * pstep + prow; p < pixels; p += (int64_t)gridDim.x * pstep) {
        size_t o = (size_t)p * C + c;
        float4 g = *reinterpret_cast<const float4 *>(dy + o);
        float4 xv = *reinterpret_cast<const float4 *>(x + o);
        float gr[4] = {g.x, g.y, g.z, g.w}, xr[4] = {xv.x, xv.y, xv.z, xv.w}, r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float pre = fmaf(xr[j], s_scale[c + j], s_shift[c + j]);
            float d = (!bn.relu || pre > 0.f) ? gr[j] : 0.f;
            float xh = (xr[j] - s_mean[c + j]) * s_istd[c + j];
            r[j] = d; a0[j] += d; a1[j] += d * xh;
        }
        *reinterpret_cast<float4 *>(dz + o) = make_float4(r[0], r[1], r[2], r[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_red[0][tid][j] = a0[j]; s_red[1][tid][j] = a1[j]; }
    __syncthreads();
    if (tid < cq4) {
        double t0[4] = {0, 0, 0, 0}, t1[4] = {0, 0, 0, 0};
        for (int i = tid; i < BT; i += cq4)
#pragma unroll
            for (int j = 0; j < 4; ++j) { t0[j] += (double)s_red[0][i][j]; t1[j] += (double)s_red[1][i][j]; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&dz_stats[tid * 4 + j], t0[j]);
            atomicAdd(&dz_stats[C + tid * 4 + j], t1[j]);
        }
    }
}

__global__ void k_bn_ema(const dpp_bn_ema_item *__restrict__ items, float alpha) {
    const dpp_bn_ema_item it = items[blockIdx.x];
    for (int c = threadIdx.x; c < it.C; c += blockDim.x) {
        double m = it.sums[c] / it.count;
        double var = it.sums[it.C + c] / it.count - m * m;
        if (var < 0.0) var = 0.0;
        float mean = (float)m, istd = (float)(1.0 / sqrt(var + (double)it.eps));
        it.mean[c] = (1.f - alpha) * it.mean[c] + alpha * mean;
        it.inv_std[c] = (1.f - alpha) * it.inv_std[c] + alpha * istd;
    }
}

int ew_grid(int64_t pixels, int C) {
    int pstep = BT / (C / 4);
    int64_t blocks = (pixels + pstep - 1) / pstep;
    int64_t cap = 148 * 8;
    return (int)(blocks < cap ? blocks : cap);
}

}  // namespace

extern "C" int dpp_bn_bwd_apply(const float *dz, const float *x, const dpp_bn_ref *bn, const double *dz_stats,
                                const float *skip, float *dx, float *dgamma, float *dbeta, double *dbias_stats,
                                int64_t pixels, int C, float param_grad_scale, void *stream) {
    DPP_CHECK_ARG(dz && x && bn && dz_stats && dx && pixels > 0);
    DPP_CHECK_ARG(C % 4 == 0 && C <= 256 && 256 % (C / 4) == 0 && bn->sums != nullptr);
    DPP_CUDA(launch_pdl(1, k_bn_bwd_apply, dim3(ew_grid(pixels, C)), dim3(BT), 0, S(stream), dz, x, *bn, dz_stats, skip, dx,
                        dgamma, dbeta, dbias_stats, pixels, C, param_grad_scale));
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

extern "C" int dpp_bn_apply(const float *x, const dpp_bn_ref *bn, float *y, int64_t pixels, int C, void *stream) {
    DPP_CHECK_ARG(x && bn && y && pixels > 0 && C % 4 == 0 && C <= 256 && 256 % (C / 4) == 0);
    k_bn_apply<<<ew_grid(pixels, C), BT, 0, S(stream)>>>(x, *bn, y, pixels, C);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

extern "C" int dpp_bn_relu_bwd_reduce(const float *dy, const float *x, const dpp_bn_ref *bn, float *dz,
                                      double *dz_stats, int64_t pixels, int C, void *stream) {
    DPP_CHECK_ARG(dy && x && bn && dz && dz_stats && pixels > 0 && C % 4 == 0 && C <= 256 && 256 % (C / 4) == 0);
    // every block ends with 2*C fp64 atomics onto the same 2*C addresses: few, long-running blocks
    int grid = ew_grid(pixels, C);
    if (grid > 148) grid = 148;
    k_bn_relu_bwd_reduce<<<grid, BT, 0, S(stream)>>>(dy, x, *bn, dz, dz_stats, pixels, C);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

extern "C" int dpp_bn_ema_update(const dpp_bn_ema_item *items, int n_layers, float alpha, void *stream) {
    DPP_CHECK_ARG(items && n_layers > 0);
    k_bn_ema<<<n_layers, 256, 0, S(stream)>>>(items, alpha);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}
