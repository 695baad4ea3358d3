// Cost (trainer/poseregnettrainer.py:92-99) and the reference's ADAM variant
// (trainer/optimizer.py:58-90) over one flat parameter arena.
#include "common.cuh"

using namespace dpp;

namespace {

__global__ void k_loss_sqerr(const float *__restrict__ out, const float *__restrict__ target, float *__restrict__ dout,
                             float *__restrict__ cost_out, int B, int D) {
    __shared__ double red[32];
    double s = 0.0;
    const float invB = 1.f / (float)B;
    for (int i = threadIdx.x; i < B * D; i += blockDim.x) {
        float d = out[i] - target[i];
        if (dout) dout[i] = 2.f * d * invB;
        s += (double)(d * d);
    }
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        cost_out[0] = (float)(t / (double)B);
    }
}

// optimizer.py:69-88 with floatX=float32 constant folding: beta1_t == 0.9f (gamma = 1-1e-8 -> 1.0f)
// One element's update, operation by operation as the reference graph evaluates it (no FMA contraction where the reference
// rounds twice).  HBM-bound: 28 bytes per parameter (w, g, m, v in; w, m, v out) - the kernel moves 16-byte vectors.
__device__ __forceinline__ void adam_one(float &w, float g, float &m, float &v, float lr, float gs, float c1, float c2) {
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    const float omb1 = 1.f - b1, omb2 = 1.f - b2;
    const float gi = g * gs;
    const float mi = __fadd_rn(__fmul_rn(b1, m), __fmul_rn(omb1, gi));
    const float vi = __fadd_rn(__fmul_rn(b2, v), __fmul_rn(omb2, __fmul_rn(gi, gi)));
    const float mh = mi / c1, vh = vi / c2;
    w = w - (lr * mh) / (sqrtf(vh) + eps);
    m = mi;
    v = vi;
}

__global__ void __launch_bounds__(256)
k_adam(float *__restrict__ w, const float *__restrict__ g, float *__restrict__ m,
       float *__restrict__ v, const float *__restrict__ hyper, int64_t n) {
    const float lr = hyper[0], t = hyper[1], gs = hyper[3];
    const float c1 = 1.f - powf(0.9f, t), c2 = 1.f - powf(0.999f, t);
    const int64_t n4 = n >> 2, stride = (int64_t)gridDim.x * blockDim.x;
    float4 *w4 = reinterpret_cast<float4 *>(w), *m4 = reinterpret_cast<float4 *>(m), *v4 = reinterpret_cast<float4 *>(v);
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 wv = w4[i], mv = m4[i], vv = v4[i];
        const float4 gv = __ldcs(g4 + i);            // the gradient is dead after this step
        adam_one(wv.x, gv.x, mv.x, vv.x, lr, gs, c1, c2);
        adam_one(wv.y, gv.y, mv.y, vv.y, lr, gs, c1, c2);
        adam_one(wv.z, gv.z, mv.z, vv.z, lr, gs, c1, c2);
        adam_one(wv.w, gv.w, mv.w, vv.w, lr, gs, c1, c2);
        w4[i] = wv; m4[i] = mv; v4[i] = vv;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        adam_one(w[i], g[i], m[i], v[i], lr, gs, c1, c2);
}

// arenas that are not 16-byte aligned (never the engine's: slots are 16-byte aligned)
__global__ void k_adam_scalar(float *__restrict__ w, const float *__restrict__ g, float *__restrict__ m,
                              float *__restrict__ v, const float *__restrict__ hyper, int64_t n) {
    const float lr = hyper[0], t = hyper[1], gs = hyper[3];
    const float c1 = 1.f - powf(0.9f, t), c2 = 1.f - powf(0.999f, t);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        adam_one(w[i], g[i], m[i], v[i], lr, gs, c1, c2);
}

__global__ void k_adam_tick(float *hyper) { hyper[1] += 1.f; }

}  // namespace

extern "C" int dpp_loss_sqerr(const float *out, const float *target, float *dout, float *cost_out, int B, int D,
                              void *stream) {
    DPP_CHECK_ARG(out && target && cost_out && B > 0 && D > 0);
    k_loss_sqerr<<<1, 1024, 0, S(stream)>>>(out, target, dout, cost_out, B, D);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

extern "C" int dpp_adam_step(float *w, const float *g, float *m, float *v, const float *hyper, int64_t n,
                             void *stream) {
    DPP_CHECK_ARG(w && g && m && v && hyper && n > 0);
    const bool aligned = ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                           reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    int64_t blocks = ((aligned ? n / 4 : n) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    if (aligned) k_adam<<<(int)blocks, 256, 0, S(stream)>>>(w, g, m, v, hyper, n);
    else k_adam_scalar<<<(int)blocks, 256, 0, S(stream)>>>(w, g, m, v, hyper, n);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

extern "C" int dpp_adam_tick(float *hyper, void *stream) {
    DPP_CHECK_ARG(hyper);
    k_adam_tick<<<1, 1, 0, S(stream)>>>(hyper);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}
