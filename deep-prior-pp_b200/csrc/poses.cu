// Random pose sampling for the PCA prior: HandDetector.sampleRandomPoses (reference util/handdetector.py:805-909,
// rot3D = False as every entry script calls it, main_nyu_posereg_embedding.py:87-88) for n poses at once.
//
// The host draws the random parameters with the reference's NumPy stream in the reference's order (:837-841) and
// passes them in; one thread computes one joint of one sampled pose.  The arithmetic follows the reference's dtype
// discipline operation by operation (float32 array arithmetic, float64 wherever a Python float / float64 array
// enters, float32 stores) with explicit round-to-nearest intrinsics so that nvcc cannot contract a*b+c into an
// FMA; cos / sin of the rotation come from the host (NumPy's libm), so results are bit-identical to the oracle.
//
// Reference pieces restated: importer.joint3DToImg / jointImgTo3D (data/importers.py:80-119, :756-793,
// :1187-1224), rotatePoint2D (data/transformations.py:71-88).
#include "common.cuh"

using namespace dpp;

namespace {

struct Cam {
    double fx, fy, ux, uy;
    int flip_y;
};

// importer.joint3DToImg on a float32 sample: s0/s2 in float32, the rest in float64, float32 store
__device__ __forceinline__ void to_img(const Cam &c, float s0, float s1, float s2, float &u, float &v, float &d) {
    if (s2 == 0.f) {
        u = (float)c.ux;
        v = (float)c.uy;
        d = 0.f;
        return;
    }
    const double q0 = (double)__fdiv_rn(s0, s2), q1 = (double)__fdiv_rn(s1, s2);
    u = (float)__dadd_rn(__dmul_rn(q0, c.fx), c.ux);
    v = c.flip_y ? (float)__dsub_rn(c.uy, __dmul_rn(q1, c.fy)) : (float)__dadd_rn(__dmul_rn(q1, c.fy), c.uy);
    d = s2;
}

// importer.jointImgTo3D: float64 expression, float32 store
__device__ __forceinline__ void to_3d(const Cam &c, float s0, float s1, float s2, float &x, float &y, float &z) {
    x = (float)__ddiv_rn(__dmul_rn(__dsub_rn((double)s0, c.ux), (double)s2), c.fx);
    y = c.flip_y ? (float)__ddiv_rn(__dmul_rn(__dsub_rn(c.uy, (double)s1), (double)s2), c.fy)
                 : (float)__ddiv_rn(__dmul_rn(__dsub_rn((double)s1, c.uy), (double)s2), c.fy);
    z = s2;
}

// joints3DToImg -> rotatePoints2D about `centre` -> jointsImgTo3D for one point
__device__ __forceinline__ void rotate_in_image(const Cam &c, float p0, float p1, float p2, float c0, float c1, double ca,
                                                double sa, float &x, float &y, float &z) {
    float u, v, d;
    to_img(c, p0, p1, p2, u, v, d);
    const float pp0 = __fsub_rn(u, c0), pp1 = __fsub_rn(v, c1);                     // float32
    const float pr0 = (float)__dsub_rn(__dmul_rn((double)pp0, ca), __dmul_rn((double)pp1, sa));
    const float pr1 = (float)__dadd_rn(__dmul_rn((double)pp0, sa), __dmul_rn((double)pp1, ca));
    to_3d(c, __fadd_rn(pr0, c0), __fadd_rn(pr1, c1), d, x, y, z);
}

__global__ void k_sample_poses(const float *__restrict__ base_poses, const float *__restrict__ base_com,
                               const float *__restrict__ base_cube, const int *__restrict__ mode,
                               const int *__restrict__ ridx, const double *__restrict__ off,
                               const double *__restrict__ sc, const double *__restrict__ cs, Cam cam,
                               float *__restrict__ new_poses, float *__restrict__ new_com,
                               float *__restrict__ new_cube, int n, int J) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * J) return;
    const int i = (int)(t / J), j = (int)(t - (long long)i * J);
    const int r = ridx[i], m = mode[i];
    const float cube[3] = {base_cube[r * 3], base_cube[r * 3 + 1], base_cube[r * 3 + 2]};
    const float com[3] = {base_com[r * 3], base_com[r * 3 + 1], base_com[r * 3 + 2]};
    const float *pp = base_poses + ((size_t)r * J + j) * 3;
    float p[3] = {pp[0], pp[1], pp[2]};
    float ncom[3] = {com[0], com[1], com[2]};
    float ncube[3] = {cube[0], cube[1], cube[2]};
    const bool moved = (m == 3 || m == 4 || m == 5);
    if (moved) {                                  // new_com = com3D + off (float64), stored float32
#pragma unroll
        for (int k = 0; k < 3; ++k) ncom[k] = (float)__dadd_rn((double)com[k], off[(size_t)i * 3 + k]);
    }
    if (m == 2) {                                 // new_cube = cube * sc: float32 array * float64 scalar -> float32
        const float s = (float)sc[i];
#pragma unroll
        for (int k = 0; k < 3; ++k) ncube[k] = __fmul_rn(cube[k], s);
    }
    const float half = (float)__ddiv_rn((double)ncube[2], 2.0);
    float o[3];
    if (m == 0 || m == 2) {                       // none / sc
#pragma unroll
        for (int k = 0; k < 3; ++k) o[k] = p[k];
    } else if (m == 3) {                          // com: pose + com3D - new_com
#pragma unroll
        for (int k = 0; k < 3; ++k) o[k] = __fsub_rn(__fadd_rn(p[k], com[k]), ncom[k]);
    } else {
        const double ca = cs[(size_t)i * 2], sa = cs[(size_t)i * 2 + 1];
        float cu, cv, cd, q[3];
        if (m == 1) {                             // rot about the projected com3D
            to_img(cam, com[0], com[1], com[2], cu, cv, cd);
#pragma unroll
            for (int k = 0; k < 3; ++k) q[k] = __fadd_rn(p[k], ncom[k]);
        } else {                                  // rot+com (4) / rot+com+sc (5): rotate about the projected new com
            to_img(cam, ncom[0], ncom[1], ncom[2], cu, cv, cd);
            const float s = (float)sc[i];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float v = __fsub_rn(__fadd_rn(p[k], com[k]), ncom[k]);
                if (m == 5) v = __fmul_rn(v, s);
                q[k] = __fadd_rn(v, com[k]);
            }
        }
        float x, y, z;
        rotate_in_image(cam, q[0], q[1], q[2], cu, cv, ca, sa, x, y, z);
        const float *sub = (m == 1) ? ncom : com;
        o[0] = __fsub_rn(x, sub[0]);
        o[1] = __fsub_rn(y, sub[1]);
        o[2] = __fsub_rn(z, sub[2]);
    }
    float *dst = new_poses + (size_t)t * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) dst[k] = __fdiv_rn(o[k], half);
    if (j == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            new_com[(size_t)i * 3 + k] = ncom[k];
            new_cube[(size_t)i * 3 + k] = ncube[k];
        }
    }
}

}  // namespace

extern "C" int dpp_sample_poses(const float *base_poses, const float *base_com, const float *base_cube, const int32_t *mode,
                                const int32_t *ridx, const double *off, const double *sc, const double *cos_sin,
                                double fx, double fy, double ux, double uy, int flip_y, float *new_poses, float *new_com,
                                float *new_cube, int n, int J, void *stream) {
    DPP_CHECK_ARG(base_poses && base_com && base_cube && mode && ridx && off && sc && cos_sin);
    DPP_CHECK_ARG(new_poses && new_com && new_cube && n >= 0 && J > 0);
    if (n == 0) return DPP_OK;
    Cam cam{fx, fy, ux, uy, flip_y};
    const long long total = (long long)n * J;
    k_sample_poses<<<cdiv(total, 256), 256, 0, S(stream)>>>(base_poses, base_com, base_cube, mode, ridx, off, sc, cos_sin,
                                                           cam, new_poses, new_com, new_cube, n, J);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}
