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
__global__ void k_adam(float *__restrict__ w, const float *__restrict__ g, float *__restrict__ m,
                       float *__restrict__ v, const float *__restrict__ hyper, int64_t n) {
    const float lr = hyper[0], t = hyper[1], gs = hyper[3];
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    const float c1 = 1.f - powf(b1, t), c2 = 1.f - powf(b2, t);
    const float omb1 = 1.f - b1, omb2 = 1.f - b2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i] * gs;
        float mi = __fadd_rn(__fmul_rn(b1, m[i]), __fmul_rn(omb1, gi));
        float vi = __fadd_rn(__fmul_rn(b2, v[i]), __fmul_rn(omb2, __fmul_rn(gi, gi)));
        float mh = mi / c1, vh = vi / c2;
        w[i] = w[i] - (lr * mh) / (sqrtf(vh) + eps);
        m[i] = mi;
        v[i] = vi;
    }
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
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_adam<<<(int)blocks, 256, 0, S(stream)>>>(w, g, m, v, hyper, n);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

extern "C" int dpp_adam_tick(float *hyper, void *stream) {
    DPP_CHECK_ARG(hyper);
    k_adam_tick<<<1, 1, 0, S(stream)>>>(hyper);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}
