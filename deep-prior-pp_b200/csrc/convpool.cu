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


// -------------------------------------------------------------------------------------------------
// ResNet stem specialisation: 5x5 'half' conv, Cin = 1, Cout = 32, 2x2 max-pool (net/resnet.py:128-133).
// fp32 FMA-bound (K = 25, one input channel: no GEMM shape worth the tensor pipe), so the kernels are built around
// FMA density.  Forward: a CTA owns an 8 x 16 region of pooled pixels; a thread owns TWO horizontally adjacent
// pooled pixels x 8 channels = 64 accumulators, keeps a sliding 2 x 8 input window in registers and runs 25 taps x
// 8 conv pixels x 8 channels of FMAs against broadcast weight vectors (64 FMAs per two 16-byte weight loads);
// BatchNorm statistics stay in registers until the end of the CTA.  The next region's input patch is loaded
// while the current one is computed (one barrier per region).
// -------------------------------------------------------------------------------------------------
constexpr int SPH = 20, SPW = 36;   // forward input patch: 8*2 + 4 rows, 16*2 + 4 columns

// tile index -> (image, tile row, tile column) without integer divisions in the per-tile path: magic multipliers
// computed once per thread (exact while tile * divisor < 2^32, i.e. for any tensor this library addresses with ints)
struct TileDiv {
    unsigned m_img, m_row;
    int per_img, per_row;
    __device__ __forceinline__ TileDiv(int tilesY, int tilesX)
        : m_img(0xFFFFFFFFu / (unsigned)(tilesY * tilesX) + 1u), m_row(0xFFFFFFFFu / (unsigned)tilesX + 1u),
          per_img(tilesY * tilesX), per_row(tilesX) {}
    __device__ __forceinline__ void operator()(int tile, int &n, int &ty, int &tx) const {
        n = per_img == 1 ? tile : (int)__umulhi((unsigned)tile, m_img);
        const int tr = tile - n * per_img;
        ty = per_row == 1 ? tr : (int)__umulhi((unsigned)tr, m_row);
        tx = tr - ty * per_row;
    }
};

// a thread's share of a region's input patch (3 of the 720 values): global loads into registers ...
__device__ __forceinline__ void stem_fetch_fwd(float (&v)[3], const float *__restrict__ x, int n, int ty0, int tx0, int H, int W) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int i = threadIdx.x + j * 256;
        const int py = i / SPW, px = i - py * SPW;
        const int yy = ty0 * 16 - 2 + py, xx = tx0 * 32 - 2 + px;
        v[j] = (i < SPH * SPW && yy >= 0 && yy < H && xx >= 0 && xx < W) ? x[((size_t)n * H + yy) * W + xx] : 0.f;
    }
}
// ... and, after the arithmetic that hides their latency, into shared memory
__device__ __forceinline__ void stem_put_fwd(float *patch, const float (&v)[3]) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int i = threadIdx.x + j * 256;
        if (i < SPH * SPW) patch[i] = v[j];
    }
}

__global__ void __launch_bounds__(256, 2)
k_stem_fwd(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
           float *__restrict__ y, uint8_t *__restrict__ argmax, double *__restrict__ stats, int N, int H, int W) {
    __shared__ __align__(16) float ws[25 * 32];
    __shared__ __align__(16) float patch[2][SPH * SPW];
    __shared__ double ssum[64];
    const int tid = threadIdx.x;
    for (int i = tid; i < 25 * 32; i += 256) ws[i] = w[i];
    if (tid < 64) ssum[tid] = 0.0;
    const int Hp = H / 2, Wp = W / 2;
    const int tilesY = (Hp + 7) / 8, tilesX = (Wp + 15) / 16;
    const int tiles = N * tilesY * tilesX;
    const TileDiv tdiv(tilesY, tilesX);
    const int pair = tid & 63, grp = tid >> 6;     // grp is warp-uniform: weight reads are broadcasts
    const int ly = pair >> 3, lxp = pair & 7;
    const int o0 = grp * 8;
    float bq[8], st1[8], st2[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { bq[q] = bias[o0 + q]; st1[q] = 0.f; st2[q] = 0.f; }
    int buf = 0;
    float pf[3];
    if ((int)blockIdx.x < tiles) {
        const int t = blockIdx.x;
        int n_, ty_, tx_;
        tdiv(t, n_, ty_, tx_);
        stem_fetch_fwd(pf, x, n_, ty_, tx_, H, W);
        stem_put_fwd(patch[0], pf);
    }
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, buf ^= 1) {
        int n, ty0, tx0;
        tdiv(tile, n, ty0, tx0);
        __syncthreads();            // patch[buf] complete; everybody is done with patch[buf ^ 1]
        const int t2 = tile + gridDim.x;
        if (t2 < tiles) {
            int n_, ty_, tx_;
            tdiv(t2, n_, ty_, tx_);
            stem_fetch_fwd(pf, x, n_, ty_, tx_, H, W);
        }
        const float *pw0 = patch[buf] + (ly * 2) * SPW + lxp * 4;
        // [pooled pixel][pool cell cy*2+cx][channel pair]: one FFMA2 (fma.rn.f32x2: two correctly rounded fp32 FMAs) per
        // pair - a 3-register FFMA issues only every second cycle per scheduler, i.e. the scalar form caps this kernel at
        // half the fp32 peak.  Same products, same order per channel: results are bit-identical to the scalar kernel.
        float2 acc[2][4][4];
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][c][q] = make_float2(0.f, 0.f);
        float rowa[8], rowb[8];     // window rows r and r + 1
        {
            const float4 a0 = *reinterpret_cast<const float4 *>(pw0), a1 = *reinterpret_cast<const float4 *>(pw0 + 4);
            rowa[0] = a0.x; rowa[1] = a0.y; rowa[2] = a0.z; rowa[3] = a0.w; rowa[4] = a1.x; rowa[5] = a1.y; rowa[6] = a1.z; rowa[7] = a1.w;
        }
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            {
                const float4 b0 = *reinterpret_cast<const float4 *>(pw0 + (r + 1) * SPW);
                const float4 b1 = *reinterpret_cast<const float4 *>(pw0 + (r + 1) * SPW + 4);
                rowb[0] = b0.x; rowb[1] = b0.y; rowb[2] = b0.z; rowb[3] = b0.w; rowb[4] = b1.x; rowb[5] = b1.y; rowb[6] = b1.z; rowb[7] = b1.w;
            }
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const float4 w0 = *reinterpret_cast<const float4 *>(&ws[(r * 5 + s) * 32 + o0]);
                const float4 w1 = *reinterpret_cast<const float4 *>(&ws[(r * 5 + s) * 32 + o0 + 4]);
                const float2 wq[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
                for (int p = 0; p < 2; ++p)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float xs = (c >> 1) ? rowb[p * 2 + (c & 1) + s] : rowa[p * 2 + (c & 1) + s];
                        const float2 xv = make_float2(xs, xs);
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[p][c][q] = __ffma2_rn(xv, wq[q], acc[p][c][q]);
                    }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) rowa[j] = rowb[j];
        }
        if (t2 < tiles) stem_put_fwd(patch[buf ^ 1], pf);
        const int ph = ty0 * 8 + ly;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int pw = tx0 * 16 + lxp * 2 + p;
            const bool valid = ph < Hp && pw < Wp;
            float v[8];
            uint32_t am[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float a4[4] = {(q & 1) ? acc[p][0][q >> 1].y : acc[p][0][q >> 1].x, (q & 1) ? acc[p][1][q >> 1].y : acc[p][1][q >> 1].x,
                                     (q & 1) ? acc[p][2][q >> 1].y : acc[p][2][q >> 1].x, (q & 1) ? acc[p][3][q >> 1].y : acc[p][3][q >> 1].x};
                float b = a4[0];
                uint32_t bi = 0;
#pragma unroll
                for (int c = 1; c < 4; ++c)
                    if (a4[c] > b) { b = a4[c]; bi = (uint32_t)c; }       // first max wins
                v[q] = b + bq[q];
                am[q] = bi;
            }
            if (valid) {
                const size_t ob = (((size_t)n * Hp + ph) * Wp + pw) * 32 + o0;
                *reinterpret_cast<float4 *>(y + ob) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4 *>(y + ob + 4) = make_float4(v[4], v[5], v[6], v[7]);
                if (argmax) {
                    uint2 pk;
                    pk.x = am[0] | (am[1] << 8) | (am[2] << 16) | (am[3] << 24);
                    pk.y = am[4] | (am[5] << 8) | (am[6] << 16) | (am[7] << 24);
                    *reinterpret_cast<uint2 *>(argmax + ob) = pk;
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) { st1[q] += v[q]; st2[q] = fmaf(v[q], v[q], st2[q]); }
            }
        }
    }
    if (stats) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const double s0 = warp_sum_d((double)st1[q]), s1 = warp_sum_d((double)st2[q]);
            if ((tid & 31) == 0) { atomicAdd(&ssum[o0 + q], s0); atomicAdd(&ssum[32 + o0 + q], s1); }
        }
        __syncthreads();
        if (tid < 64) atomicAdd(&stats[tid], ssum[tid]);
    }
}

// stem weight / bias gradients.  Lane = output channel, warp = 8 of the tile's 64 pooled pixels: a thread reads
// its own (pixel, channel) gradient and arg-max cell straight from global memory (128-byte rows per warp), turns
// the cell into the patch offset of its 5x5 window and accumulates all 25 taps in registers (one shared-memory
// read with an immediate offset + one FMA per tap).  The 8 warps meet in shared memory at the end of the CTA:
// one atomicAdd per (tap, channel) per CTA.
constexpr int SP = 20;   // backward input patch edge: 8*2 + 5 - 1

__device__ __forceinline__ void stem_fetch_bwd(float (&v)[2], const float *__restrict__ x, int n, int ty0, int tx0, int H, int W) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int i = threadIdx.x + j * 256;
        const int py = i / SP, px = i - py * SP;
        const int yy = ty0 * 16 - 2 + py, xx = tx0 * 16 - 2 + px;
        v[j] = (i < SP * SP && yy >= 0 && yy < H && xx >= 0 && xx < W) ? x[((size_t)n * H + yy) * W + xx] : 0.f;
    }
}
__device__ __forceinline__ void stem_put_bwd(float *patch, const float (&v)[2]) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int i = threadIdx.x + j * 256;
        if (i < SP * SP) patch[i] = v[j];
    }
}

__global__ void __launch_bounds__(256)
k_stem_bwd_w(const float *__restrict__ x, const uint8_t *__restrict__ argmax, const float *__restrict__ dy,
             float *__restrict__ dw, float *__restrict__ db, int N, int H, int W) {
    __shared__ float patch[2][SP * SP];
    __shared__ float red[8][26][32];
    const int tid = threadIdx.x;
    const int o = tid & 31, wq = tid >> 5;
    const int Hp = H / 2, Wp = W / 2;
    const int tilesY = (Hp + 7) / 8, tilesX = (Wp + 7) / 8;
    const int tiles = N * tilesY * tilesX;
    const TileDiv tdiv(tilesY, tilesX);
    float acc[25], accb = 0.f;
#pragma unroll
    for (int t = 0; t < 25; ++t) acc[t] = 0.f;
    int buf = 0;
    float g[8], gn[8], pf[2];
    int base[8], cn[8];
    // this warp's pooled row ty0*8 + wq, columns tx0*8 .. +7 of a tile: gradient and arg-max cell per (pixel, lane = channel)
    auto fetch = [&](int tile) {
        int n, ty0, tx0;
        tdiv(tile, n, ty0, tx0);
        const int ph = ty0 * 8 + wq;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int pw = tx0 * 8 + j;
            const bool valid = ph < Hp && pw < Wp;
            const size_t ob = (((size_t)n * Hp + (valid ? ph : 0)) * Wp + (valid ? pw : 0)) * 32 + o;
            gn[j] = valid ? dy[ob] : 0.f;
            cn[j] = valid ? argmax[ob] : 0;
        }
    };
    if ((int)blockIdx.x < tiles) {
        const int t = blockIdx.x;
        int n_, ty_, tx_;
        tdiv(t, n_, ty_, tx_);
        stem_fetch_bwd(pf, x, n_, ty_, tx_, H, W);
        stem_put_bwd(patch[0], pf);
        fetch(t);
    }
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, buf ^= 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            g[j] = gn[j];
            base[j] = (wq * 2 + (cn[j] >> 1)) * SP + j * 2 + (cn[j] & 1);
        }
        __syncthreads();            // patch[buf] complete; everybody is done with patch[buf ^ 1]
        const int t2 = tile + gridDim.x;
        if (t2 < tiles) {
            int n_, ty_, tx_;
            tdiv(t2, n_, ty_, tx_);
            stem_fetch_bwd(pf, x, n_, ty_, tx_, H, W);
            fetch(t2);
        }
        const float *pb = patch[buf];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float *pp = pb + base[j];
#pragma unroll
            for (int t = 0; t < 25; ++t) acc[t] = fmaf(g[j], pp[(t / 5) * SP + (t % 5)], acc[t]);
            accb += g[j];
        }
        if (t2 < tiles) stem_put_bwd(patch[buf ^ 1], pf);
    }
#pragma unroll
    for (int t = 0; t < 25; ++t) red[wq][t][o] = acc[t];
    red[wq][25][o] = accb;
    __syncthreads();
    for (int i = tid; i < 26 * 32; i += 256) {
        const int t = i >> 5, oo = i & 31;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][t][oo];
        if (t < 25) atomicAdd(&dw[t * 32 + oo], s); else atomicAdd(&db[oo], s);
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

// convpool8.cu: specialised kernels for the 8-filter 'valid' towers (default; DPP_CONVPOOL_FAST=0 switches them off)
int dpp_convpool8_try(const float *x, const float *w, const float *bias, float *y, uint8_t *argmax, double *stats, int N,
                      int H, int W, int Cin, int Cout, int k, int pad, int pool, int relu, void *stream);

extern "C" int dpp_convpool_fwd(const float *x, const float *w, const float *bias, float *y, uint8_t *argmax,
                                double *stats, int N, int H, int W, int Cin, int Cout, int k, int pad, int pool,
                                int relu, void *stream) {
    DPP_CHECK_ARG(x && w && bias && y && N > 0 && Cin > 0 && Cout > 0 && k > 0 && pool >= 1 && pool <= 8);
    {
        const int rc8 = dpp_convpool8_try(x, w, bias, y, argmax, stats, N, H, W, Cin, Cout, k, pad, pool, relu, stream);
        if (rc8 != DPP_ENOTSUP) return rc8;
    }
    CPDims d;
    DPP_CHECK_ARG(make_dims(d, N, H, W, Cin, Cout, k, pad, pool) == 0);
    size_t smem = sizeof(float) * ((size_t)k * k * Cin * Cout + (size_t)d.patch * d.patch * Cin + 2 * Cout);
    DPP_CHECK_ARG(smem <= 200 * 1024);
    DPP_CUDA(cudaFuncSetAttribute(k_convpool_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int tiles = N * d.tilesY * d.tilesX;
    if (k == 5 && pool == 2 && Cin == 1 && Cout == 32 && pad == 2 && !relu && H % 2 == 0 && W % 2 == 0) {
        const int t2 = N * ((d.Hp + 7) / 8) * ((d.Wp + 15) / 16);     // 8 x 16 pooled regions
        const int g2 = t2 < 148 * 2 ? t2 : 148 * 2;
        k_stem_fwd<<<g2, 256, 0, S(stream)>>>(x, w, bias, y, argmax, stats, N, H, W);
        DPP_LAUNCH_CHECK();
        return DPP_OK;
    }
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
    if (k == 5 && pool == 2 && Cin == 1 && Cout == 32 && pad == 2 && !relu && !dx && H % 2 == 0 && W % 2 == 0) {
        int g2 = tiles < 148 * 4 ? tiles : 148 * 4;
        k_stem_bwd_w<<<g2, 256, 0, S(stream)>>>(x, argmax, dy, dw, db, N, H, W);
        DPP_LAUNCH_CHECK();
        return DPP_OK;
    }
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
