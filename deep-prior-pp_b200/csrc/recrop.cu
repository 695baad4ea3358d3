// Fused hand crop of the inference cascade: window of the depth frame -> zero padding -> z-threshold ->
// nearest-neighbour resize -> centred paste -> CoM normalisation (-> the 1/2 and 1/4 centre crops ScaleNet
// takes), one CTA per output crop.
//
// Replaces, per frame, HandDetector.getCrop (reference util/handdetector.py:260-296), resizeCrop = cv2.resize
// INTER_NEAREST (:336-351), the paste of cropArea3D (:467-476), the normalisations of refineCoM (:640-647) and
// RealtimeHandposePipeline.detect (util/realtimehandposepipeline.py:327-332), refineCoM's centre crops
// (:656-667) and estimatePose's mirroring (:346-349).  The window geometry (comToBounds, :204-226) is host
// work in fp64 (util/handdetector.py mirror) and arrives as one dpp_crop_rec per output crop.
//
// Index rule (cv2 4.13.0 resizeNN, oracle/cascade.py::resize_nn_indices): source = min(floor(x * ifx), n-1) with
// ifx = 1. / (n_dst / n_src) in fp64, computed on the host so that no device division decides an index; the
// product x * ifx is an explicit round-to-nearest multiply.
//
// HBM-bound: 4 B read per sampled pixel + 4 B written (+ 5/16 for the two centre crops).  The sampled pixels of
// one output row lie in one contiguous span of a frame row, so a warp's gathers touch neighbouring sectors.
#include "common.cuh"

using namespace dpp;

namespace {

constexpr int RC_THREADS = 256;
constexpr int RC_FILL = -1;     // outside the pasted patch: the filler value (getNDValue)
constexpr int RC_PAD = -2;      // inside the window but outside the frame: getCrop's zero padding

__device__ __forceinline__ float recrop_px(const float *__restrict__ frame, int sx, int sy, int Wf,
                                           const dpp_crop_rec &r) {
    float v;
    if (sx == RC_FILL || sy == RC_FILL) {
        v = r.fill;
    } else {
        v = (sx >= 0 && sy >= 0) ? __ldg(frame + (size_t)sy * Wf + sx) : 0.f;
        // getCrop, handdetector.py:291-295
        bool m1 = (v < r.zstart) && (v != 0.f);
        bool m2 = (v > r.zend) && (v != 0.f);
        if (m1) v = r.zstart;
        if (m2) v = 0.f;
    }
    if (r.flags & 1) {
        if (v == 0.f) v = r.hi;
        if (r.flags & 2) {       // refineCoM clamps; the pipeline's crop.clip() result is discarded
            if (v >= r.hi) v = r.hi;
            if (v <= r.lo) v = r.lo;
        }
        v = __fdiv_rn(__fsub_rn(v, r.comz), r.half);
    }
    return v;
}

__global__ void __launch_bounds__(RC_THREADS)
k_recrop(const float *__restrict__ frames, const dpp_crop_rec *__restrict__ recs, float *__restrict__ out0,
         float *__restrict__ out1, float *__restrict__ out2, int Hf, int Wf, int H, int W) {
    extern __shared__ int rc_tab[];
    int *sxt = rc_tab;          // [W] source column of every output column (or RC_FILL / RC_PAD)
    int *syt = rc_tab + W;      // [H]
    const dpp_crop_rec r = recs[blockIdx.x];
    const int tid = threadIdx.x;

    for (int i = tid; i < W + H; i += RC_THREADS) {
        const bool isx = i < W;
        const int d = isx ? i - r.px : (i - W) - r.py;
        const int nd = isx ? r.rw : r.rh, ns = isx ? r.wb : r.hb, s0 = isx ? r.xstart : r.ystart;
        const int lim = isx ? Wf : Hf;
        int s = RC_FILL;
        if (d >= 0 && d < nd) {
            s = (int)floor(__dmul_rn((double)d, isx ? r.ifx : r.ify));
            if (s > ns - 1) s = ns - 1;
            s += s0;
            if (s < 0 || s >= lim) s = RC_PAD;
        }
        rc_tab[i] = s;
    }
    __syncthreads();

    const float *frame = frames + (size_t)r.src_index * Hf * Wf;
    const int qw = W >> 2;
    const bool mirror = (r.flags & 4) != 0;
    float *o0 = out0 + (size_t)blockIdx.x * H * W;
    const int H2 = H >> 1, W2 = W >> 1, H4 = H >> 2, W4 = W >> 2;
    const int y1 = (H - H2) >> 1, x1 = (W - W2) >> 1;    // refineCoM: int(H/2 - (H//2)/2)
    const int y2 = (H - H4) >> 1, x2 = (W - W4) >> 1;
    for (int q = tid; q < H * qw; q += RC_THREADS) {
        const int y = q / qw, x0 = (q - y * qw) << 2;
        const int sy = syt[y];
        float4 v;
        v.x = recrop_px(frame, sxt[x0 + 0], sy, Wf, r);
        v.y = recrop_px(frame, sxt[x0 + 1], sy, Wf, r);
        v.z = recrop_px(frame, sxt[x0 + 2], sy, Wf, r);
        v.w = recrop_px(frame, sxt[x0 + 3], sy, Wf, r);
        int xo = x0;
        if (mirror) {
            float t = v.x; v.x = v.w; v.w = t;
            t = v.y; v.y = v.z; v.z = t;
            xo = W - 4 - x0;
        }
        *reinterpret_cast<float4 *>(o0 + (size_t)y * W + xo) = v;
        if (out1 != nullptr && y >= y1 && y < y1 + H2 && xo >= x1 && xo < x1 + W2)
            *reinterpret_cast<float4 *>(out1 + ((size_t)blockIdx.x * H2 + (y - y1)) * W2 + (xo - x1)) = v;
        if (out2 != nullptr && y >= y2 && y < y2 + H4 && xo >= x2 && xo < x2 + W4)
            *reinterpret_cast<float4 *>(out2 + ((size_t)blockIdx.x * H4 + (y - y2)) * W4 + (xo - x2)) = v;
    }
}

// ---- pose error metrics -----------------------------------------------------------------------------------
// Reference util/handpose_evaluation.py:92-181 (getMeanError / getMaxError / get*OverSeq / getJointErrorOverSeq: all are
// reductions of sqrt(square(gt - joints).sum(axis=2))) and trainer/poseregnettrainer.py:123-125 (errors_avg /
// errors_max).  One warp per frame: Euclidean error of every joint, then the frame's nan-mean and nan-max.
__global__ void k_joint_errors(const float *__restrict__ pred, const float *__restrict__ gt, float *__restrict__ err,
                               float *__restrict__ frame_mean, float *__restrict__ frame_max, int J) {
    __shared__ float s_sum[32], s_max[32];
    __shared__ int s_cnt[32];
    const int f = blockIdx.x, tid = threadIdx.x;
    float sum = 0.f, mx = 0.f;
    int cnt = 0;
    for (int j = tid; j < J; j += blockDim.x) {
        const float *p = pred + ((size_t)f * J + j) * 3, *g = gt + ((size_t)f * J + j) * 3;
        float dx = __fsub_rn(g[0], p[0]), dy = __fsub_rn(g[1], p[1]), dz = __fsub_rn(g[2], p[2]);
        // numpy: sqrt(((gt - joints)**2).sum(axis=2)): squares rounded one by one, summed left to right
        float e = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        if (err != nullptr) err[(size_t)f * J + j] = e;
        if (e == e) {            // numpy.nanmean / nanmax: joints without annotation are NaN
            sum += e;
            mx = fmaxf(mx, e);
            ++cnt;
        }
    }
    sum = warp_sum(sum);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) { s_sum[tid >> 5] = sum; s_max[tid >> 5] = mx; s_cnt[tid >> 5] = cnt; }
    __syncthreads();
    if (tid == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        float a = 0.f, m = 0.f;
        int c = 0;
        for (int i = 0; i < nw; ++i) { a += s_sum[i]; m = fmaxf(m, s_max[i]); c += s_cnt[i]; }
        if (frame_mean != nullptr) frame_mean[f] = a / (float)c;      // 0/0 = NaN like numpy.nanmean of all-NaN
        if (frame_max != nullptr) frame_max[f] = c > 0 ? m : __int_as_float(0x7fc00000);
    }
}

}  // namespace

extern "C" int dpp_recrop_fwd(const float *frames, const dpp_crop_rec *recs, float *out0, float *out1, float *out2,
                              int n_out, int Hf, int Wf, int H, int W, void *stream) {
    DPP_CHECK_ARG(frames && recs && out0 && n_out >= 0 && Hf > 0 && Wf > 0 && H > 0 && W > 0);
    DPP_CHECK_ARG(W % 32 == 0 && H % 8 == 0 && W <= 4096 && H <= 4096);
    if (n_out == 0) return DPP_OK;
    k_recrop<<<n_out, RC_THREADS, sizeof(int) * (W + H), S(stream)>>>(frames, recs, out0, out1, out2, Hf, Wf, H, W);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

extern "C" int dpp_joint_errors(const float *pred, const float *gt, float *err, float *frame_mean, float *frame_max,
                                int n_frames, int J, void *stream) {
    DPP_CHECK_ARG(pred && gt && n_frames >= 0 && J > 0);
    if (n_frames == 0) return DPP_OK;
    k_joint_errors<<<n_frames, 32, 0, S(stream)>>>(pred, gt, err, frame_mean, frame_max, J);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}
