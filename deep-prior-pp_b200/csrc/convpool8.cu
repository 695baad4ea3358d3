// Specialised ConvPoolLayer forward for the 8-filter 'valid' towers of ScaleNet (net/scalenet.py:55-107) and
// PoseRegNet (net/poseregnet.py:60-78): conv (k = 5 | 3, Cin = 1 | 8) -> max-pool (4 | 2 | 1) -> +bias -> ReLU.
//
// Why: at batch 1024 the generic k_convpool_fwd spends 3.8 of the cascade's 11.1 ms here (profiles/README.md): with
// Cout = 8 only 64 of its 256 threads have work, and its tap / channel / pool loops have run-time bounds.  Here the
// loop bounds are template parameters (fully unrolled FMAs out of registers), the pooled tile is 16 x 16 so every
// thread owns one pooled pixel x 8 channels, and the weights are read as broadcast shared-memory vectors.
//
// STATUS: default path since round 2 (DPP_CONVPOOL_FAST=0 selects the generic kernel).  Verified on a B200:
// tests/test_gpu_convpool8.py - outputs and arg-max cells bit for bit equal to the generic kernel for every tower
// shape; cascade batch of 1024 frames 11.15 -> 7.98 ms (profiles/README.md, round 2).  tests/test_host_convpool8.py
// runs the shared per-thread code (convpool8.cuh) on the host against an independent reference.
#include <stdlib.h>
#include "common.cuh"
#include "convpool8.cuh"

using namespace dpp;

namespace {

constexpr int TP8 = 16;                 // pooled tile edge: 256 threads = 256 pooled pixels

template <int K, int CIN, int POOL>
__global__ void __launch_bounds__(TP8 * TP8)
k_convpool8_fwd(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                float *__restrict__ y, uint8_t *__restrict__ argmax, int N, int H, int W, int Hp, int Wp, int tilesY,
                int tilesX, int relu) {
    constexpr int P = TP8 * POOL + K - 1;
    extern __shared__ __align__(16) float sm8[];
    float *ws = sm8;                                 // K*K*CIN*8
    float *patch = ws + K * K * CIN * 8;             // P*P*CIN
    for (int i = threadIdx.x; i < K * K * CIN * 8; i += TP8 * TP8) ws[i] = w[i];
    const int ly = threadIdx.x / TP8, lx = threadIdx.x % TP8;
    const int tiles = N * tilesY * tilesX;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int n = tile / (tilesY * tilesX), tr = tile % (tilesY * tilesX);
        const int ty0 = tr / tilesX, tx0 = tr % tilesX;
        const int y0 = ty0 * TP8 * POOL, x0 = tx0 * TP8 * POOL;     // 'valid': no padding
        __syncthreads();                             // previous tile's patch is no longer read (also orders ws)
        for (int i = threadIdx.x; i < P * P * CIN; i += TP8 * TP8) {
            const int c = i % CIN, pp = i / CIN;
            const int px = pp % P, py = pp / P;
            const int yy = y0 + py, xx = x0 + px;
            patch[i] = (yy < H && xx < W) ? x[(((size_t)n * H + yy) * W + xx) * CIN + c] : 0.f;
        }
        __syncthreads();
        const int ph = ty0 * TP8 + ly, pw = tx0 * TP8 + lx;
        if (ph < Hp && pw < Wp) {
            float best[8];
            uint8_t bidx[8];
            convpool8_pixel<K, CIN, POOL>(patch, P, ws, ly, lx, best, bidx);
            convpool8_store(best, bidx, bias, relu, y, argmax, (((size_t)n * Hp + ph) * Wp + pw) * 8);
        }
    }
}

template <int K, int CIN, int POOL>
int launch8(const float *x, const float *w, const float *bias, float *y, uint8_t *argmax, int N, int H, int W, int relu,
            cudaStream_t st) {
    constexpr int P = TP8 * POOL + K - 1;
    const int Hp = (H - K + 1) / POOL, Wp = (W - K + 1) / POOL;
    if (Hp <= 0 || Wp <= 0) return DPP_ENOTSUP;
    const int tilesY = (Hp + TP8 - 1) / TP8, tilesX = (Wp + TP8 - 1) / TP8;
    const int tiles = N * tilesY * tilesX;
    const size_t smem = sizeof(float) * (K * K * CIN * 8 + P * P * CIN);
    const int grid = tiles < 148 * 8 ? tiles : 148 * 8;
    k_convpool8_fwd<K, CIN, POOL><<<grid, TP8 * TP8, smem, st>>>(x, w, bias, y, argmax, N, H, W, Hp, Wp, tilesY, tilesX,
                                                              relu);
    return DPP_OK;
}

}  // namespace

// Returns DPP_ENOTSUP when the configuration (or the DPP_CONVPOOL_FAST switch) does not select this kernel.
int dpp_convpool8_try(const float *x, const float *w, const float *bias, float *y, uint8_t *argmax, double *stats, int N,
                      int H, int W, int Cin, int Cout, int k, int pad, int pool, int relu, void *stream) {
    const char *e = getenv("DPP_CONVPOOL_FAST");
    if (e != nullptr && e[0] == '0') return DPP_ENOTSUP;
    if (Cout != 8 || pad != 0 || stats != nullptr) return DPP_ENOTSUP;
    cudaStream_t st = S(stream);
    int rc = DPP_ENOTSUP;
    if (k == 5 && Cin == 1 && pool == 4) rc = launch8<5, 1, 4>(x, w, bias, y, argmax, N, H, W, relu, st);
    else if (k == 5 && Cin == 1 && pool == 2) rc = launch8<5, 1, 2>(x, w, bias, y, argmax, N, H, W, relu, st);
    else if (k == 5 && Cin == 8 && pool == 2) rc = launch8<5, 8, 2>(x, w, bias, y, argmax, N, H, W, relu, st);
    else if (k == 5 && Cin == 8 && pool == 1) rc = launch8<5, 8, 1>(x, w, bias, y, argmax, N, H, W, relu, st);
    else if (k == 3 && Cin == 8 && pool == 1) rc = launch8<3, 8, 1>(x, w, bias, y, argmax, N, H, W, relu, st);
    if (rc == DPP_OK) DPP_LAUNCH_CHECK();
    return rc;
}
