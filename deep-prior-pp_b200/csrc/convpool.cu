// ConvPoolLayer: conv (true convolution, 'half' or 'valid') -> max-pool (floor) -> +bias -> act.
// Reference: net/convpoollayer.py:251-282 (theano conv2d + pool_2d(ignore_border=True) + bias).
// Used for the ResNet stem (1->32, 5x5 'half', pool 2; net/resnet.py:128-133) and for the small
// PoseRegNet / ScaleNet towers (8 filters, 'valid'; net/poseregnet.py:60-78).
//
// Cin is 1 or 8 here, so this is a direct convolution out of a shared-memory input patch, not a
// GEMM: the stem has K = 25 and is bound by the 64 KiB/crop input read + pooled output write.
// Persistent CTAs loop over 8x8 pooled tiles so BN statistics are reduced per CTA, not per tile.
#include "common.cuh"

using namespace dpp;

namespace {

constexpr int TP = 8;           // pooled tile edge
constexpr int CG = 8;           // output channels per thread
constexpr int CP_THREADS = 256;

struct CPDims {
    int N, H, W, Cin, Cout, k, pad, pool, Hp, Wp, tilesY, tilesX, patch;  // patch edge
};

__device__ __forceinline__ void load_patch(float *patch, const float *__restrict__ x, const CPDims &d, int n,
                                           int ty0, int tx0) {
    // patch[(py*P + px)*Cin + c] for conv-input rows ty0*pool-pad .. (+P)
    const int P = d.patch;
    const int y0 = ty0 * TP * d.pool - d.pad, x0 = tx0 * TP * d.pool - d.pad;
    const int total = P * P * d.Cin;
    for (int i = threadIdx.x; i < total; i += CP_THREADS) {
        int c = i % d.Cin, pp = i / d.Cin;
        int px = pp % P, py = pp / P;
        int yy = y0 + py, xx = x0 + px;
        float v = 0.f;
        if (yy >= 0 && yy < d.H && xx >= 0 && xx < d.W) v = x[(((size_t)n * d.H + yy) * d.W + xx) * d.Cin + c];
        patch[i] = v;
    }
}

__global__ void __launch_bounds__(CP_THREADS)
k_convpool_fwd(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
               float *__restrict__ y, uint8_t *__restrict__ argmax, double *__restrict__ stats, CPDims d, int relu) {
    extern __shared__ __align__(16) float sm[];
    const int K = d.k * d.k * d.Cin;
    float *ws = sm;                        // K*Cout
    float *patch = ws + K * d.Cout;        // P*P*Cin
    float *ssum = patch + d.patch * d.patch * d.Cin;   // 2*Cout (per-CTA stats)
    for (int i = threadIdx.x; i < K * d.Cout; i += CP_THREADS) ws[i] = w[i];
    for (int i = threadIdx.x; i < 2 * d.Cout; i += CP_THREADS) ssum[i] = 0.f;

    const int groups = (d.Cout + CG - 1) / CG;
    const int items = TP * TP * groups;
    const int tiles = d.N * d.tilesY * d.tilesX;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        int n = tile / (d.tilesY * d.tilesX), tr = tile % (d.tilesY * d.tilesX);
        int ty0 = tr / d.tilesX, tx0 = tr % d.tilesX;
        __syncthreads();
        load_patch(patch, x, d, n, ty0, tx0);
        __syncthreads();
        for (int it = threadIdx.x; it < items; it += CP_THREADS) {
            int pix = it % (TP * TP), grp = it / (TP * TP);
            int ly = pix / TP, lx = pix % TP;
            int ph = ty0 * TP + ly, pw = tx0 * TP + lx;
            if (ph >= d.Hp || pw >= d.Wp) continue;
            int o0 = grp * CG;
            float best[CG];
            int bestc = 0;
            uint8_t bidx[CG];
#pragma unroll
            for (int q = 0; q < CG; ++q) { best[q] = -INFINITY; bidx[q] = 0; }
            (void)bestc;
            for (int cy = 0; cy < d.pool; ++cy)
                for (int cx = 0; cx < d.pool; ++cx) {
                    float acc[CG];
#pragma unroll
                    for (int q = 0; q < CG; ++q) acc[q] = 0.f;
                    const int by = ly * d.pool + cy, bx = lx * d.pool + cx;
                    for (int r = 0; r < d.k; ++r)
                        for (int s = 0; s < d.k; ++s) {
                            const float *pp = patch + ((by + r) * d.patch + (bx + s)) * d.Cin;
                            const float *wp = ws + ((r * d.k + s) * d.Cin) * d.Cout + o0;
                            for (int c = 0; c < d.Cin; ++c) {
                                float xv = pp[c];
#pragma unroll
                                for (int q = 0; q < CG; ++q)
                                    if (o0 + q < d.Cout) acc[q] = fmaf(xv, wp[c * d.Cout + q], acc[q]);
                            }
                        }
                    uint8_t cell = (uint8_t)(cy * d.pool + cx);
#pragma unroll
                    for (int q = 0; q < CG; ++q)
                        if (acc[q] > best[q]) { best[q] = acc[q]; bidx[q] = cell; }   // first max wins
                }
            size_t ob = (((size_t)n * d.Hp + ph) * d.Wp + pw) * d.Cout + o0;
#pragma unroll
            for (int q = 0; q < CG; ++q) {
                if (o0 + q >= d.Cout) break;
                float v = best[q] + bias[o0 + q];
                if (relu) v = fmaxf(v, 0.f);
                y[ob + q] = v;
                if (argmax) argmax[ob + q] = bidx[q];
                if (stats) {
                    atomicAdd(&ssum[o0 + q], v);
                    atomicAdd(&ssum[d.Cout + o0 + q], v * v);
                }
            }
        }
        if (stats) {   // flush per tile in fp64 to keep the fp32 partial sums short
            __syncthreads();
            for (int i = threadIdx.x; i < 2 * d.Cout; i += CP_THREADS) {
                atomicAdd(&stats[i], (double)ssum[i]);
                ssum[i] = 0.f;
            }
        }
    }
}

// dW/db accumulation.  One work item = (pooled pixel, channel group); a warp holds 32 pixels of
// one group, so the per-tap products are reduced with shuffles and added to shared accumulators.
__global__ void __launch_bounds__(CP_THREADS)
k_convpool_bwd_w(const float *__restrict__ x, const float *__restrict__ y, const uint8_t *__restrict__ argmax,
                 const float *__restrict__ dy, float *__restrict__ dw, float *__restrict__ db, CPDims d, int relu) {
    extern __shared__ __align__(16) float sm[];
    const int K = d.k * d.k * d.Cin;
    float *dws = sm;                       // K*Cout
    float *dbs = dws + K * d.Cout;         // Cout
    float *patch = dbs + d.Cout;
    for (int i = threadIdx.x; i < K * d.Cout + d.Cout; i += CP_THREADS) dws[i] = 0.f;

    const int groups = (d.Cout + CG - 1) / CG;
    const int items = TP * TP * groups;          // multiple of 32
    const int tiles = d.N * d.tilesY * d.tilesX;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        int n = tile / (d.tilesY * d.tilesX), tr = tile % (d.tilesY * d.tilesX);
        int ty0 = tr / d.tilesX, tx0 = tr % d.tilesX;
        __syncthreads();
        load_patch(patch, x, d, n, ty0, tx0);
        __syncthreads();
        for (int it0 = (threadIdx.x & ~31); it0 < items; it0 += CP_THREADS) {
            int it = it0 + (threadIdx.x & 31);
            int pix = it % (TP * TP), grp = it / (TP * TP);     // 32 lanes: same grp (TP*TP = 64)
            int ly = pix / TP, lx = pix % TP;
            int ph = ty0 * TP + ly, pw = tx0 * TP + lx;
            bool valid = (ph < d.Hp && pw < d.Wp);
            int o0 = grp * CG;
            float g[CG];
            int cell[CG];
            size_t ob = (((size_t)n * d.Hp + (valid ? ph : 0)) * d.Wp + (valid ? pw : 0)) * d.Cout + o0;
#pragma unroll
            for (int q = 0; q < CG; ++q) {
                g[q] = 0.f;
                cell[q] = 0;
                if (valid && o0 + q < d.Cout) {
                    float gv = dy[ob + q];
                    if (relu && !(y[ob + q] > 0.f)) gv = 0.f;
                    g[q] = gv;
                    cell[q] = argmax[ob + q];
                }
            }
#pragma unroll
            for (int q = 0; q < CG; ++q) {
                float s = warp_sum(g[q]);
                if ((threadIdx.x & 31) == 0 && o0 + q < d.Cout) atomicAdd(&dbs[o0 + q], s);
            }
            for (int r = 0; r < d.k; ++r)
                for (int s = 0; s < d.k; ++s)
                    for (int c = 0; c < d.Cin; ++c) {
#pragma unroll
                        for (int q = 0; q < CG; ++q) {
                            int cy = cell[q] / d.pool, cx = cell[q] % d.pool;
                            float xv = patch[((ly * d.pool + cy + r) * d.patch + (lx * d.pool + cx + s)) * d.Cin + c];
                            float p = warp_sum(g[q] * xv);
                            if ((threadIdx.x & 31) == 0 && o0 + q < d.Cout)
                                atomicAdd(&dws[((r * d.k + s) * d.Cin + c) * d.Cout + o0 + q], p);
                        }
                    }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * d.Cout; i += CP_THREADS) atomicAdd(&dw[i], dws[i]);
    for (int i = threadIdx.x; i < d.Cout; i += CP_THREADS) atomicAdd(&db[i], dbs[i]);
}

// dx for the towers whose input is itself trainable-network output (PoseRegNet/ScaleNet layers
// 2,3).  One thread per (pooled pixel, out channel) scatters into dx with atomics (tiny nets).
__global__ void k_convpool_bwd_x(const float *__restrict__ w, const float *__restrict__ y,
                                 const uint8_t *__restrict__ argmax, const float *__restrict__ dy,
                                 float *__restrict__ dx, CPDims d, int relu) {
    size_t total = (size_t)d.N * d.Hp * d.Wp * d.Cout;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int o = (int)(i % d.Cout);
        size_t p = i / d.Cout;
        int pw = (int)(p % d.Wp);
        int ph = (int)((p / d.Wp) % d.Hp);
        int n = (int)(p / ((size_t)d.Wp * d.Hp));
        float g = dy[i];
        if (relu && !(y[i] > 0.f)) g = 0.f;
        if (g == 0.f) continue;
        int cell = argmax[i];
        int hc = ph * d.pool + cell / d.pool, wc = pw * d.pool + cell % d.pool;
        for (int r = 0; r < d.k; ++r)
            for (int s = 0; s < d.k; ++s) {
                int yy = hc - d.pad + r, xx = wc - d.pad + s;
                if (yy < 0 || yy >= d.H || xx < 0 || xx >= d.W) continue;
                for (int c = 0; c < d.Cin; ++c)
                    atomicAdd(&dx[(((size_t)n * d.H + yy) * d.W + xx) * d.Cin + c],
                              g * w[((r * d.k + s) * d.Cin + c) * d.Cout + o]);
            }
    }
}

int make_dims(CPDims &d, int N, int H, int W, int Cin, int Cout, int k, int pad, int pool) {
    d.N = N; d.H = H; d.W = W; d.Cin = Cin; d.Cout = Cout; d.k = k; d.pad = pad; d.pool = pool;
    int Hc = H + 2 * pad - k + 1, Wc = W + 2 * pad - k + 1;
    d.Hp = Hc / pool; d.Wp = Wc / pool;
    d.tilesY = (d.Hp + TP - 1) / TP; d.tilesX = (d.Wp + TP - 1) / TP;
    d.patch = TP * pool + k - 1;
    return (d.Hp > 0 && d.Wp > 0) ? 0 : -1;
}

}  // namespace

extern "C" int dpp_convpool_fwd(const float *x, const float *w, const float *bias, float *y, uint8_t *argmax,
                                double *stats, int N, int H, int W, int Cin, int Cout, int k, int pad, int pool,
                                int relu, void *stream) {
    DPP_CHECK_ARG(x && w && bias && y && N > 0 && Cin > 0 && Cout > 0 && k > 0 && pool >= 1 && pool <= 8);
    CPDims d;
    DPP_CHECK_ARG(make_dims(d, N, H, W, Cin, Cout, k, pad, pool) == 0);
    size_t smem = sizeof(float) * ((size_t)k * k * Cin * Cout + (size_t)d.patch * d.patch * Cin + 2 * Cout);
    DPP_CHECK_ARG(smem <= 200 * 1024);
    DPP_CUDA(cudaFuncSetAttribute(k_convpool_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int tiles = N * d.tilesY * d.tilesX;
    int grid = tiles < 148 * 4 ? tiles : 148 * 4;
    k_convpool_fwd<<<grid, CP_THREADS, smem, S(stream)>>>(x, w, bias, y, argmax, stats, d, relu);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

extern "C" int dpp_convpool_bwd(const float *x, const float *w, const float *y, const uint8_t *argmax, const float *dy,
                                float *dw, float *db, float *dx, int N, int H, int W, int Cin, int Cout, int k,
                                int pad, int pool, int relu, void *stream) {
    DPP_CHECK_ARG(x && w && y && argmax && dy && dw && db && N > 0 && pool >= 1 && pool <= 8);
    CPDims d;
    DPP_CHECK_ARG(make_dims(d, N, H, W, Cin, Cout, k, pad, pool) == 0);
    size_t smem = sizeof(float) * ((size_t)k * k * Cin * Cout + Cout + (size_t)d.patch * d.patch * Cin);
    DPP_CHECK_ARG(smem <= 200 * 1024);
    DPP_CUDA(cudaFuncSetAttribute(k_convpool_bwd_w, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int tiles = N * d.tilesY * d.tilesX;
    int grid = tiles < 148 * 2 ? tiles : 148 * 2;
    k_convpool_bwd_w<<<grid, CP_THREADS, smem, S(stream)>>>(x, y, argmax, dy, dw, db, d, relu);
    DPP_LAUNCH_CHECK();
    if (dx) {
        DPP_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)N * H * W * Cin, S(stream)));
        size_t total = (size_t)N * d.Hp * d.Wp * Cout;
        int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
        k_convpool_bwd_x<<<blocks, 256, 0, S(stream)>>>(w, y, argmax, dy, dx, d, relu);
        DPP_LAUNCH_CHECK();
    }
    return DPP_OK;
}
