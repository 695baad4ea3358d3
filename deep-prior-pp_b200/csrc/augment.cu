// Fused depth-crop augmentation: de-normalise -> (none | affine NN | perspective NN +
// z-threshold) -> CoM re-normalise, one CTA per output crop.
//
// Replaces the pixel work of NetTrainer.augmentCrop (reference trainer/nettrainer.py:948-995)
// and the cv2.warpAffine / cv2.warpPerspective(INTER_NEAREST, BORDER_CONSTANT 0) calls of
// HandDetector.rotateHand / recropHand (reference util/handdetector.py:730-738, :791-801).
// Index rules are those of oracle/augment.py (cv2 4.13.0): every fp64 operation below is an
// explicit round-to-nearest intrinsic so that nvcc cannot contract a*b+c into an FMA - the
// crop matrices put whole rows on exact .5 ties and one ulp decides the source pixel.
//
// HBM-bound: 64 KiB read + 64 KiB written per crop (128x128 fp32); the source crop is staged
// in shared memory once (coalesced float4 loads), the gather then runs out of shared memory.
#include "common.cuh"

using namespace dpp;

namespace {

constexpr int AUG_THREADS = 512;

__device__ __forceinline__ float finalize_px(float v, float premax, const dpp_aug_rec &r) {
    if (r.mode & 16) return v;   // raw: warp only (HandDetector.rotateHand/recropHand called directly)
    // nettrainer.py:990-995, applied in the reference's order
    if (v == premax) v = r.bg;
    if (v == 0.f) v = r.bg;
    if (v >= r.bg) v = r.bg;
    if (v <= r.lo) v = r.lo;
    return __fdiv_rn(__fsub_rn(v, r.comz_new), r.half_new);
}

__global__ void __launch_bounds__(AUG_THREADS, 1)
k_augment(const float *__restrict__ crops, const dpp_aug_rec *__restrict__ recs, float *__restrict__ out,
          int H, int W) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *src = reinterpret_cast<float *>(smem_raw);              // H*W de-normalised source
    float *dst = src + H * W;                                      // H*W staging (mode 2)
    double *rowtab = reinterpret_cast<double *>(dst + H * W);      // 3*H (mode 2)
    int *itab = reinterpret_cast<int *>(rowtab + 3 * H);           // 2*W + 2*H (mode 1)
    __shared__ float red[AUG_THREADS / 32];
    __shared__ float s_premax;

    const dpp_aug_rec r = recs[blockIdx.x];
    const int tid = threadIdx.x;
    const int npx = H * W;
    const float *g = crops + (size_t)r.src_index * npx;
    float *o = out + (size_t)blockIdx.x * npx;

    // 1. de-normalise (nettrainer.py:951: img*(cube_z/2) + com_z, two fp32 roundings) + max
    float mx = -INFINITY;
    for (int i = tid * 4; i < npx; i += AUG_THREADS * 4) {
        float4 v = *reinterpret_cast<const float4 *>(g + i);
        v.x = __fadd_rn(__fmul_rn(v.x, r.half_old), r.comz_old);
        v.y = __fadd_rn(__fmul_rn(v.y, r.half_old), r.comz_old);
        v.z = __fadd_rn(__fmul_rn(v.z, r.half_old), r.comz_old);
        v.w = __fadd_rn(__fmul_rn(v.w, r.half_old), r.comz_old);
        *reinterpret_cast<float4 *>(src + i) = v;
        mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if ((tid & 31) == 0) red[tid >> 5] = mx;

    // 2. per-sample tables
    const int wmode = r.mode & 15;
    if (wmode == 2) {
        if (tid < 3) {                       // running fp64 row sums r(0)=c, r(y+1)=r(y)+b
            double b = r.m[tid * 3 + 1], v = r.m[tid * 3 + 2];
            for (int y = 0; y < H; ++y) {
                rowtab[tid * H + y] = v;
                v = __dadd_rn(v, b);
            }
        }
    } else if (wmode == 1) {
        // AB_BITS = 10 fixed point (cv2 warpAffine): adelta/bdelta per x, X0/Y0 per y
        for (int i = tid; i < W; i += AUG_THREADS) {
            itab[i] = (int)rint(__dmul_rn(__dmul_rn(r.m[0], (double)i), 1024.0));
            itab[W + i] = (int)rint(__dmul_rn(__dmul_rn(r.m[3], (double)i), 1024.0));
        }
        for (int i = tid; i < H; i += AUG_THREADS) {
            itab[2 * W + i] = (int)rint(__dmul_rn(__dadd_rn(__dmul_rn(r.m[1], (double)i), r.m[2]), 1024.0)) + 512;
            itab[2 * W + H + i] = (int)rint(__dmul_rn(__dadd_rn(__dmul_rn(r.m[4], (double)i), r.m[5]), 1024.0)) + 512;
        }
    }
    __syncthreads();
    if (tid == 0) {
        float m = red[0];
        for (int i = 1; i < AUG_THREADS / 32; ++i) m = fmaxf(m, red[i]);
        s_premax = m;
    }
    __syncthreads();
    const float premax = s_premax;

    if (wmode == 0) {
        for (int i = tid; i < npx; i += AUG_THREADS) o[i] = finalize_px(src[i], premax, r);
    } else if (wmode == 1) {
        for (int i = tid; i < npx; i += AUG_THREADS) {
            int y = i / W, x = i - y * W;
            int X = (itab[2 * W + y] + itab[x]) >> 10;
            int Y = (itab[2 * W + H + y] + itab[W + x]) >> 10;
            float v = (X >= 0 && X < W && Y >= 0 && Y < H) ? src[Y * W + X] : 0.f;
            o[i] = finalize_px(v, premax, r);
        }
    } else {
        // cv2 4.13 warpPerspective NN: 4 fp64 lane accumulators per row, see oracle.
        const double a0 = r.m[0], a1 = r.m[3], a2 = r.m[6];
        const double s0 = __dmul_rn(4.0, a0), s1 = __dmul_rn(4.0, a1), s2 = __dmul_rn(4.0, a2);
        const double wmax = (double)(W - 1), hmax = (double)(H - 1);
        for (int t = tid; t < H * 4; t += AUG_THREADS) {
            int y = t >> 2, j = t & 3;
            double nx = __dadd_rn(__dmul_rn((double)j, a0), rowtab[y]);
            double ny = __dadd_rn(__dmul_rn((double)j, a1), rowtab[H + y]);
            double nw = __dadd_rn(__dmul_rn((double)j, a2), rowtab[2 * H + y]);
            for (int x = j; x < W; x += 4) {
                double sx = __ddiv_rn(nx, nw), sy = __ddiv_rn(ny, nw);
                float v = 0.f;
                if (sx >= 0.0 && sx <= wmax && sy >= 0.0 && sy <= hmax) {
                    int X = (int)floor(__dadd_rn(sx, 0.5));
                    int Y = (int)floor(__dadd_rn(sy, 0.5));
                    v = src[Y * W + X];
                }
                // recropHand (handdetector.py:793-801)
                if (fabsf(__fsub_rn(v, 32000.f)) <= 0.32000001f) v = 0.f;
                bool m1 = (v < r.zstart) && (v != 0.f);
                bool m2 = (v > r.zend) && (v != 0.f);
                if (m1) v = r.zstart;
                if (m2) v = 0.f;
                dst[y * W + x] = finalize_px(v, premax, r);
                nx = __dadd_rn(nx, s0);
                ny = __dadd_rn(ny, s1);
                nw = __dadd_rn(nw, s2);
            }
        }
        __syncthreads();
        for (int i = tid * 4; i < npx; i += AUG_THREADS * 4)
            *reinterpret_cast<float4 *>(o + i) = *reinterpret_cast<const float4 *>(dst + i);
    }
}

}  // namespace

extern "C" int dpp_augment_fwd(const float *crops, const dpp_aug_rec *recs, float *out, int n_out, int H, int W,
                               void *stream) {
    DPP_CHECK_ARG(crops && recs && out && n_out >= 0 && H > 0 && W > 0);
    DPP_CHECK_ARG(W % 4 == 0 && (H * W) % 4 == 0 && H * W <= 128 * 128);
    if (n_out == 0) return DPP_OK;
    size_t smem = sizeof(float) * 2 * H * W + sizeof(double) * 3 * H + sizeof(int) * (2 * W + 2 * H);
    static bool attr_done = false;
    if (!attr_done) {
        DPP_CUDA(cudaFuncSetAttribute(k_augment, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_done = true;
    }
    k_augment<<<n_out, AUG_THREADS, smem, S(stream)>>>(crops, recs, out, H, W);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}
