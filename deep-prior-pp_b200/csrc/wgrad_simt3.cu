// Backward weights of the 3x3, stride-1, 'half'-padded, Cin = Cout = C in {16, 32} ConvLayers (conv2 of the first two
// residual stages, reference net/resnet.py:60-75 through T.grad, trainer/poseregnettrainer.py:110-111) as an fp32 FMA kernel:
//   dw[(r*3+s)*C + ci][co] += sum_p a(p + (r-1, s-1))[ci] * dy[p][co],   a = ReLU(BN(x)) (zero outside the image),  db[co] += sum_p dy
// As a GEMM this is M = 9C = 144 / 288, N = C = 16 / 32, K = pixels: tcgen05 runs it at a fraction of its rate (the
// instruction count is set by K/8 and an N = 16 MMA still costs 33-42 cycles; the tensor-core version in wgrad_tc_mn.cu
// needs 64 / 33 us per layer for 0.6 GFLOP), whereas the FMA pipes do 0.6 GFLOP in 8 us - and the nine taps of a pixel
// share their operand loads when the accumulators live in registers:
//   a thread owns 4 input channels x 4 output channels x 9 taps = 144 accumulators and walks every PL-th pixel of a tile
//   (PL = 16 pixel lanes for C = 16, 4 for C = 32): 10 x LDS.128 (9 activation quads + 1 gradient quad) per 144 FMAs.
// Tiles = 8 image rows (+ halo) of one image: raw activations and gradients arrive with cp.async into a double buffer
// (halo = zero fill), BatchNorm + ReLU is applied in place once per element, then the FMA loop runs while the next tile
// loads.  A CTA works through a contiguous, cost-balanced range of (layer, image, row block) items and flushes its
// accumulators (cross-lane sum through shared memory, one atomicAdd per weight) only when the layer changes.
#include "common.cuh"
#include <vector>

using namespace dpp;

namespace dpp {
struct W3Layer {
    const float *x; const float *dy; float *dw; float *db;
    dpp_bn_ref in_bn; int has_in_bn;
    int N, H, W, C;
    int rblocks;       // ceil(H / 8)
    int item0;         // first item of this layer in the global item numbering
    unsigned wp_magic, w_magic;   // i / (W + 2) = (i * wp_magic) >> 16 and i / W = (i * w_magic) >> 16 over a tile's pixel indices
    int cout;                     // 1x1 layers (k_wgrad1): Cout; there H = pixels, W = 1, C = Cin
};
}

namespace {

constexpr int R3 = 8;            // image rows per tile
constexpr int W3_THREADS = 256;

__device__ __forceinline__ void cp16(void *dst, const void *src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const uint32_t n = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

struct W3Args {
    const W3Layer *layers;
    int n_layers;
    const int *cta_begin;     // [gridDim.x + 1] item ranges (cost-balanced on the host)
    int buf_floats;           // floats per staging buffer (x tile + dy tile of the largest layer)
};

extern __shared__ __align__(16) float smem3[];     // all accesses go through this symbol (offsets, not pointers) so that they
                                                   // compile to LDS / STS rather than generic loads

// issue the loads of one item into a staging buffer: x rows [h0-1, h0+R3] x cols [-1, W] x C (zero fill outside), dy rows [h0, h0+R3)
template <int C>
__device__ __forceinline__ void stage_item(const W3Layer &L, int n, int rb, int bufo) {
    float *buf = smem3 + bufo;
    const int W = L.W, H = L.H, Wp = W + 2;
    constexpr int q = C >> 2;      // 16-byte pieces per pixel (compile-time: the index arithmetic below is shifts, not divisions)
    const int h0 = rb * R3;
    const int xpieces = (R3 + 2) * Wp * q;
    for (int i = threadIdx.x; i < xpieces; i += W3_THREADS) {
        const int piece = i % q, pix = i / q;
        const int trow = (int)(((unsigned)pix * L.wp_magic) >> 16);
        const int px = pix - trow * Wp - 1, py = trow + h0 - 1;
        const bool ok = px >= 0 && px < W && py >= 0 && py < H;
        const float *src = L.x + (((size_t)n * H + (ok ? py : 0)) * W + (ok ? px : 0)) * C + piece * 4;
        cp16(buf + (size_t)i * 4, src, ok);
    }
    float *dyb = buf + (R3 + 2) * Wp * C;
    const int dpieces = R3 * W * q;
    for (int i = threadIdx.x; i < dpieces; i += W3_THREADS) {
        const int pix = i / q;
        const int trow = (int)(((unsigned)pix * L.w_magic) >> 16);
        const int py = trow + h0;
        const bool ok = py < H;
        const float *src = L.dy + (((size_t)n * H + (ok ? py : 0)) * W) * C + (size_t)(i - trow * W * q) * 4;
        cp16(dyb + (size_t)i * 4, src, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// a = ReLU(x * scale + shift) in place, halo stays zero
template <int C>
__device__ __forceinline__ void bn_in_place(const W3Layer &L, int rb, int bufo, const float *s_scale, const float *s_shift) {
    float *buf = smem3 + bufo;
    const int W = L.W, H = L.H, Wp = W + 2;
    constexpr int q = C >> 2;
    const int h0 = rb * R3;
    const int xpieces = (R3 + 2) * Wp * q;
    const bool relu = L.in_bn.relu != 0;
    for (int i = threadIdx.x; i < xpieces; i += W3_THREADS) {
        const int piece = i % q, pix = i / q;
        const int trow = (int)(((unsigned)pix * L.wp_magic) >> 16);
        const int px = pix - trow * Wp - 1, py = trow + h0 - 1;
        if (px >= 0 && px < W && py >= 0 && py < H) {
            float4 v = *reinterpret_cast<float4 *>(buf + (size_t)i * 4);
            const int c = piece * 4;
            v.x = fmaf(v.x, s_scale[c], s_shift[c]); v.y = fmaf(v.y, s_scale[c + 1], s_shift[c + 1]);
            v.z = fmaf(v.z, s_scale[c + 2], s_shift[c + 2]); v.w = fmaf(v.w, s_scale[c + 3], s_shift[c + 3]);
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            *reinterpret_cast<float4 *>(buf + (size_t)i * 4) = v;
        }
    }
}

template <int C>
__global__ void __launch_bounds__(W3_THREADS, 1)
k_wgrad3(const W3Args a) {
    constexpr int G = C / 4;                 // channel quads
    constexpr int PL = W3_THREADS / (G * G); // pixel lanes
    __shared__ float s_scale[32], s_shift[32];
    const int tid = threadIdx.x;
    const int combo = tid % (G * G), ps = tid / (G * G);
    const int gi = combo / G, go = combo % G;
    const int bufsz = a.buf_floats;          // staging buffer b starts at float offset b * bufsz

    // accumulators as float2 pairs over the input channel: acc[t][ip][j] = {dw(ci = 2 ip, co = j), dw(ci = 2 ip + 1, co = j)}.
    // One FFMA2 (fma.rn.f32x2, two correctly rounded fp32 FMAs) per pair: a 3-register FFMA issues every second cycle
    // per scheduler on this architecture, so the scalar form caps an FMA-bound kernel at half the fp32 peak.
    float2 acc[9][2][4];
    float dbp[4];
    auto zero_acc = [&]() {
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[t][i][j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) dbp[j] = 0.f;
    };
    // sum over the pixel lanes through shared memory (one tap at a time), one atomicAdd per weight
    auto flush = [&](const W3Layer &L, int scro) {
        float *scratch = smem3 + scro;
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 9; ++t) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4 *>(scratch + ((size_t)ps * (G * G) + combo) * 16 + i * 4) =
                    (i & 1) ? make_float4(acc[t][i >> 1][0].y, acc[t][i >> 1][1].y, acc[t][i >> 1][2].y, acc[t][i >> 1][3].y)
                            : make_float4(acc[t][i >> 1][0].x, acc[t][i >> 1][1].x, acc[t][i >> 1][2].x, acc[t][i >> 1][3].x);
            __syncthreads();
            for (int o = tid; o < G * G * 16; o += W3_THREADS) {
                float s = 0.f;
#pragma unroll
                for (int l = 0; l < PL; ++l) s += scratch[(size_t)l * (G * G * 16) + o];
                const int cb = o >> 4, e = o & 15;
                const int ci = (cb / G) * 4 + (e >> 2), co = (cb % G) * 4 + (e & 3);
                atomicAdd(L.dw + ((size_t)t * C + ci) * C + co, s);
            }
            __syncthreads();
        }
        if (L.db != nullptr) {
            if (gi == 0) *reinterpret_cast<float4 *>(scratch + ((size_t)ps * G + go) * 4) = make_float4(dbp[0], dbp[1], dbp[2], dbp[3]);
            __syncthreads();
            if (tid < C) {
                float s = 0.f;
#pragma unroll
                for (int l = 0; l < PL; ++l) s += scratch[(size_t)l * C + tid];
                atomicAdd(L.db + tid, s);
            }
            __syncthreads();
        }
    };

    const int it_begin = a.cta_begin[blockIdx.x], it_end = a.cta_begin[blockIdx.x + 1];
    if (it_begin >= it_end) return;
    // item -> (layer, image, row block); items of a layer are contiguous
    int li = 0;
    while (li + 1 < a.n_layers && a.layers[li + 1].item0 <= it_begin) ++li;
    W3Layer L = a.layers[li];
    auto load_coef = [&]() {
        if (tid < C) {
            float sc = 1.f, sh = 0.f;
            if (L.has_in_bn) bn_scale_shift(L.in_bn, tid, C, sc, sh);
            s_scale[tid] = sc; s_shift[tid] = sh;
        }
    };
    load_coef();
    zero_acc();
    int buf = 0;
    {
        const int r = it_begin - L.item0;
        stage_item<C>(L, r / L.rblocks, r % L.rblocks, 0);
    }
    for (int it = it_begin; it < it_end; ++it, buf ^= 1) {
        const int r = it - L.item0;
        const int rb = r % L.rblocks;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                    // tile `it` has landed; everybody is done with the other buffer
        if (L.has_in_bn) bn_in_place<C>(L, rb, buf * bufsz, s_scale, s_shift);
        // next item: same layer -> prefetch now; a new layer starts after the flush below
        const int nxt = it + 1;
        const bool more = nxt < it_end;
        const bool same = more && (li + 1 >= a.n_layers || a.layers[li + 1].item0 > nxt);
        if (same) {
            const int r2 = nxt - L.item0;
            stage_item<C>(L, r2 / L.rblocks, r2 % L.rblocks, (buf ^ 1) * bufsz);
        }
        __syncthreads();                    // normalised tile visible
        {
            // A pixel lane walks a contiguous run of the tile's pixels, row segment by row segment, with a sliding
            // window of three activation columns (3 rows x 4 channels each) in registers: per pixel 3 new activation
            // quads + 1 gradient quad from shared memory (the 9 taps of neighbouring pixels share 6 of their 9 quads)
            // - LDS.128 returns 512 bytes per warp, so at 10 loads per pixel the shared-memory pipe, not the FMA pipe,
            // set the pace.  The column roles rotate through a 3-way unrolled loop instead of register moves.
            const int W = L.W, Wp = W + 2, H = L.H;
            const int xb = buf * bufsz + gi * 4, dyb = buf * bufsz + (R3 + 2) * Wp * C + go * 4;      // float offsets into smem3
            const int rows = (H - rb * R3) < R3 ? (H - rb * R3) : R3;
            const int npix = rows * W;
            const int run = (R3 * W + PL - 1) / PL;
            int p = ps * run;
            const int p_end = p + run < npix ? p + run : npix;
            int py = (int)(((unsigned)p * L.w_magic) >> 16), px = p - py * W;
            float2 c0[3][2], c1[3][2], c2[3][2];      // a column: 3 rows x 4 channels as two channel pairs
            auto ldcol = [&](float2 (&c)[3][2], int base) {       // base = haloed column, tile row py (= image row py - 1)
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float4 v = *reinterpret_cast<const float4 *>(smem3 + base + r * Wp * C);
                    c[r][0] = make_float2(v.x, v.y); c[r][1] = make_float2(v.z, v.w);
                }
            };
            auto fma_px = [&](const float2 (&l)[3][2], const float2 (&m)[3][2], const float2 (&rr)[3][2], int dp) {
                const float4 d4 = *reinterpret_cast<const float4 *>(smem3 + dp);
                const float2 d[4] = {make_float2(d4.x, d4.x), make_float2(d4.y, d4.y), make_float2(d4.z, d4.z), make_float2(d4.w, d4.w)};
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            acc[r * 3 + 0][i][j] = __ffma2_rn(l[r][i], d[j], acc[r * 3 + 0][i][j]);
                            acc[r * 3 + 1][i][j] = __ffma2_rn(m[r][i], d[j], acc[r * 3 + 1][i][j]);
                            acc[r * 3 + 2][i][j] = __ffma2_rn(rr[r][i], d[j], acc[r * 3 + 2][i][j]);
                        }
                if (gi == 0) { dbp[0] += d4.x; dbp[1] += d4.y; dbp[2] += d4.z; dbp[3] += d4.w; }
            };
#pragma unroll 1
            while (p < p_end) {
                int seg = W - px;                                  // pixels of this run in image row py
                if (seg > p_end - p) seg = p_end - p;
                const int xr = xb + (py * Wp + px) * C;            // haloed column px = image column px - 1
                const int dr = dyb + p * C;
                ldcol(c0, xr);
                ldcol(c1, xr + C);
                int k = 0;
#pragma unroll 1
                for (; k + 3 <= seg; k += 3) {
                    ldcol(c2, xr + (k + 2) * C); fma_px(c0, c1, c2, dr + k * C);
                    ldcol(c0, xr + (k + 3) * C); fma_px(c1, c2, c0, dr + (k + 1) * C);
                    ldcol(c1, xr + (k + 4) * C); fma_px(c2, c0, c1, dr + (k + 2) * C);
                }
                if (k < seg) { ldcol(c2, xr + (k + 2) * C); fma_px(c0, c1, c2, dr + k * C); ++k; }
                if (k < seg) { ldcol(c0, xr + (k + 2) * C); fma_px(c1, c2, c0, dr + k * C); ++k; }
                p += seg; px += seg;
                if (px >= W) { px = 0; ++py; }
            }
        }
        if (more && !same) {                // layer boundary inside this CTA's range
            flush(L, (buf ^ 1) * bufsz);
            ++li;
            L = a.layers[li];
            load_coef();
            zero_acc();
            const int r2 = nxt - L.item0;
            stage_item<C>(L, r2 / L.rblocks, r2 % L.rblocks, (buf ^ 1) * bufsz);
        }
    }
    flush(L, buf * bufsz);     // buf was flipped after the last item: this is the buffer NOT read last (the flush syncs first anyway)
}

// ---------------------------------------------------------------------------------------------------------------------
// 1x1 stride-1 layers of the first residual stage (64 -> 16 and 16 -> 64 channels at 32 x 32, batch 128: 131072 pixels):
//   dw[ci][co] += sum_p a[p][ci] * dy[p][co],  a = ReLU(BN(x)),  db[co] += sum_p dy[p][co]
// a tall-skinny product (M x N = 16 x 64 or 64 x 16, K = pixels) that tcgen05 runs at N = 16 or M = 16 of 128 useful rows
// (30 - 35 us per layer in wgrad_tc_mn.cu); here a thread owns an 8 x 8 block of dw in registers and walks every 16th
// pixel of a 256-pixel tile: 4 x LDS.128 per 64 FMAs.  Same tile pipeline as k_wgrad3: cp.async double buffer,
// BatchNorm + ReLU in place, a contiguous cost-balanced range of (layer, tile) items per CTA, one flush per layer.
// W3Layer is reused: H = number of pixels, W = 1, C = Cin, rblocks = tiles, `cout` = Cout.
constexpr int P1 = 256;          // pixels per tile

template <int CI, int CO>
__device__ __forceinline__ void stage_item1(const W3Layer &L, int tile, int bufo) {
    float *buf = smem3 + bufo;
    const int npix = L.H - tile * P1 < P1 ? L.H - tile * P1 : P1;
    const float *xs = L.x + (size_t)tile * P1 * CI, *ds = L.dy + (size_t)tile * P1 * CO;
    for (int i = threadIdx.x; i < P1 * CI / 4; i += W3_THREADS) cp16(buf + i * 4, xs + (size_t)i * 4, i * 4 < npix * CI);
    float *dyb = buf + P1 * CI;
    for (int i = threadIdx.x; i < P1 * CO / 4; i += W3_THREADS) cp16(dyb + i * 4, ds + (size_t)i * 4, i * 4 < npix * CO);
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int CI, int CO>
__global__ void __launch_bounds__(W3_THREADS, 1)
k_wgrad1(const W3Args a) {
    constexpr int GI = CI / 8, GO = CO / 8, COMBOS = GI * GO, PL = W3_THREADS / COMBOS;
    static_assert(COMBOS * PL == W3_THREADS && PL * COMBOS * 64 <= P1 * (CI + CO), "flush scratch must fit a staging buffer");
    __shared__ float s_scale[CI], s_shift[CI];
    const int tid = threadIdx.x;
    const int combo = tid % COMBOS, ps = tid / COMBOS;
    const int gi = combo / GO, go = combo % GO;
    const int bufsz = a.buf_floats;
    float2 acc[4][8];       // pairs over the input channel: acc[ip][j] = {dw(2 ip, j), dw(2 ip + 1, j)}, one FFMA2 each
    float dbp[8];
    auto zero_acc = [&]() {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i >> 1][j] = make_float2(0.f, 0.f);
            dbp[i] = 0.f;
        }
    };
    auto flush = [&](const W3Layer &L, int scro) {
        float *scratch = smem3 + scro;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float *d = scratch + ((size_t)ps * COMBOS + combo) * 64 + i * 8;
            const float2 *ar = acc[i >> 1];
            *reinterpret_cast<float4 *>(d) = (i & 1) ? make_float4(ar[0].y, ar[1].y, ar[2].y, ar[3].y) : make_float4(ar[0].x, ar[1].x, ar[2].x, ar[3].x);
            *reinterpret_cast<float4 *>(d + 4) = (i & 1) ? make_float4(ar[4].y, ar[5].y, ar[6].y, ar[7].y) : make_float4(ar[4].x, ar[5].x, ar[6].x, ar[7].x);
        }
        __syncthreads();
        for (int o = tid; o < COMBOS * 64; o += W3_THREADS) {
            float s = 0.f;
#pragma unroll
            for (int l = 0; l < PL; ++l) s += scratch[(size_t)l * (COMBOS * 64) + o];
            const int cb = o >> 6, e = o & 63;
            const int ci = (cb / GO) * 8 + (e >> 3), co = (cb % GO) * 8 + (e & 7);
            atomicAdd(L.dw + (size_t)ci * CO + co, s);
        }
        __syncthreads();
        if (L.db != nullptr) {
            if (gi == 0) {
                float *d = scratch + ((size_t)ps * GO + go) * 8;
                *reinterpret_cast<float4 *>(d) = make_float4(dbp[0], dbp[1], dbp[2], dbp[3]);
                *reinterpret_cast<float4 *>(d + 4) = make_float4(dbp[4], dbp[5], dbp[6], dbp[7]);
            }
            __syncthreads();
            if (tid < CO) {
                float s = 0.f;
#pragma unroll
                for (int l = 0; l < PL; ++l) s += scratch[(size_t)l * CO + tid];
                atomicAdd(L.db + tid, s);
            }
            __syncthreads();
        }
    };
    const int it_begin = a.cta_begin[blockIdx.x], it_end = a.cta_begin[blockIdx.x + 1];
    if (it_begin >= it_end) return;
    int li = 0;
    while (li + 1 < a.n_layers && a.layers[li + 1].item0 <= it_begin) ++li;
    W3Layer L = a.layers[li];
    auto load_coef = [&]() {
        if (tid < CI) {
            float sc = 1.f, sh = 0.f;
            if (L.has_in_bn) bn_scale_shift(L.in_bn, tid, CI, sc, sh);
            s_scale[tid] = sc; s_shift[tid] = sh;
        }
    };
    load_coef();
    zero_acc();
    int buf = 0;
    stage_item1<CI, CO>(L, it_begin - L.item0, 0);
    for (int it = it_begin; it < it_end; ++it, buf ^= 1) {
        const int tile = it - L.item0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const int npix = L.H - tile * P1 < P1 ? L.H - tile * P1 : P1;
        if (L.has_in_bn) {
            float *xb = smem3 + buf * bufsz;
            const bool relu = L.in_bn.relu != 0;
            for (int i = tid; i < npix * CI / 4; i += W3_THREADS) {
                float4 v = *reinterpret_cast<float4 *>(xb + i * 4);
                const int c = (i * 4) % CI;
                v.x = fmaf(v.x, s_scale[c], s_shift[c]); v.y = fmaf(v.y, s_scale[c + 1], s_shift[c + 1]);
                v.z = fmaf(v.z, s_scale[c + 2], s_shift[c + 2]); v.w = fmaf(v.w, s_scale[c + 3], s_shift[c + 3]);
                if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                *reinterpret_cast<float4 *>(xb + i * 4) = v;
            }
        }
        const int nxt = it + 1;
        const bool more = nxt < it_end;
        const bool same = more && (li + 1 >= a.n_layers || a.layers[li + 1].item0 > nxt);
        if (same) stage_item1<CI, CO>(L, nxt - L.item0, (buf ^ 1) * bufsz);
        __syncthreads();
        {
            const int xo = buf * bufsz + gi * 8, dyo = buf * bufsz + P1 * CI + go * 8;
#pragma unroll 2
            for (int p = ps; p < npix; p += PL) {
                const float4 x0 = *reinterpret_cast<const float4 *>(smem3 + xo + p * CI);
                const float4 x1 = *reinterpret_cast<const float4 *>(smem3 + xo + p * CI + 4);
                const float4 d0 = *reinterpret_cast<const float4 *>(smem3 + dyo + p * CO);
                const float4 d1 = *reinterpret_cast<const float4 *>(smem3 + dyo + p * CO + 4);
                const float2 xv[4] = {make_float2(x0.x, x0.y), make_float2(x0.z, x0.w), make_float2(x1.x, x1.y), make_float2(x1.z, x1.w)};
                const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 dd = make_float2(dv[j], dv[j]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i][j] = __ffma2_rn(xv[i], dd, acc[i][j]);
                }
                if (gi == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) dbp[j] += dv[j];
                }
            }
        }
        if (more && !same) {
            flush(L, (buf ^ 1) * bufsz);
            ++li;
            L = a.layers[li];
            load_coef();
            zero_acc();
            stage_item1<CI, CO>(L, nxt - L.item0, (buf ^ 1) * bufsz);
        }
    }
    flush(L, buf * bufsz);
}

struct W3Group {
    W3Args args16, args32, args1a, args1b;        // 3x3 C = 16 / 32; 1x1 64 -> 16 / 16 -> 64
    int grid16, grid32, smem16, smem32, grid1a, grid1b, smem1a, smem1b;
    std::vector<void *> allocs;
};

// 1x1 layers with (Cin, Cout) = (CI, CO): items = tiles of 256 pixels, equal cost
template <int CI, int CO>
int build1(const std::vector<dpp::W3Layer> &all, W3Args &args, int &grid, int &smem, std::vector<void *> &allocs) {
    std::vector<dpp::W3Layer> ls;
    for (const dpp::W3Layer &l : all) if (l.W == 1 && l.C == CI && l.cout == CO) ls.push_back(l);
    grid = 0; smem = 0;
    if (ls.empty()) return 0;
    int items = 0;
    for (dpp::W3Layer &l : ls) {
        l.rblocks = (l.H + P1 - 1) / P1;
        l.item0 = items;
        items += l.rblocks;
    }
    const int bufmax = P1 * (CI + CO);
    smem = 2 * bufmax * (int)sizeof(float);
    grid = items < 148 ? items : 148;
    std::vector<int> begin(grid + 1, 0);
    for (int b = 0; b <= grid; ++b) begin[b] = (int)((long long)items * b / grid);
    void *dl = nullptr, *db = nullptr;
    if (cudaMalloc(&dl, ls.size() * sizeof(dpp::W3Layer)) != cudaSuccess || cudaMalloc(&db, begin.size() * sizeof(int)) != cudaSuccess ||
        cudaMemcpy(dl, ls.data(), ls.size() * sizeof(dpp::W3Layer), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(db, begin.data(), begin.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess)
        return -1;
    allocs.push_back(dl); allocs.push_back(db);
    args.layers = reinterpret_cast<const dpp::W3Layer *>(dl);
    args.n_layers = (int)ls.size();
    args.cta_begin = reinterpret_cast<const int *>(db);
    args.buf_floats = bufmax;
    if (cudaFuncSetAttribute(k_wgrad1<CI, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
    return 0;
}

template <int C>
int build(const std::vector<dpp::W3Layer> &all, W3Args &args, int &grid, int &smem, std::vector<void *> &allocs) {
    std::vector<dpp::W3Layer> ls;
    for (const dpp::W3Layer &l : all) if (l.C == C && l.W > 1) ls.push_back(l);
    grid = 0; smem = 0;
    if (ls.empty()) return 0;
    // items and their cost (pixels of a full tile x C^2); ranges of equal cost per CTA
    int items = 0, bufmax = 0;
    std::vector<double> cost;
    for (dpp::W3Layer &l : ls) {
        l.wp_magic = 65536u / (unsigned)(l.W + 2) + 1u;
        l.w_magic = 65536u / (unsigned)l.W + 1u;
        for (unsigned i = 0; i < (unsigned)((R3 + 2) * (l.W + 2)); ++i)
            if (((i * l.wp_magic) >> 16) != i / (unsigned)(l.W + 2) || ((i * l.w_magic) >> 16) != i / (unsigned)l.W) return -2;
        l.rblocks = (l.H + R3 - 1) / R3;
        l.item0 = items;
        const int n = l.N * l.rblocks;
        for (int i = 0; i < n; ++i) {
            const int rb = i % l.rblocks;
            const int rows = (l.H - rb * R3) < R3 ? (l.H - rb * R3) : R3;
            cost.push_back((double)rows * l.W * C * C + 4000.0);      // + staging / barrier overhead per item
        }
        items += n;
        const int bf = ((R3 + 2) * (l.W + 2) + R3 * l.W) * C;
        if (bf > bufmax) bufmax = bf;
    }
    if (bufmax < W3_THREADS * 16) bufmax = W3_THREADS * 16;           // the flush scratch needs 256 x 16 floats
    smem = 2 * bufmax * (int)sizeof(float);
    if (smem > 200 * 1024) return -2;
    grid = items < 148 ? items : 148;
    double total = 0;
    for (double c : cost) total += c;
    std::vector<int> begin(grid + 1, 0);
    {
        double run = 0;
        int b = 1;
        for (int i = 0; i < items && b < grid; ++i) {
            run += cost[i];
            while (b < grid && run >= total * b / grid) begin[b++] = i + 1;
        }
        for (; b < grid; ++b) begin[b] = items;
        begin[grid] = items;
    }
    void *dl = nullptr, *db = nullptr;
    if (cudaMalloc(&dl, ls.size() * sizeof(dpp::W3Layer)) != cudaSuccess || cudaMalloc(&db, begin.size() * sizeof(int)) != cudaSuccess ||
        cudaMemcpy(dl, ls.data(), ls.size() * sizeof(dpp::W3Layer), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(db, begin.data(), begin.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess)
        return -1;
    allocs.push_back(dl); allocs.push_back(db);
    args.layers = reinterpret_cast<const dpp::W3Layer *>(dl);
    args.n_layers = (int)ls.size();
    args.cta_begin = reinterpret_cast<const int *>(db);
    args.buf_floats = bufmax;
    if (cudaFuncSetAttribute(k_wgrad3<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
    return 0;
}

}  // namespace

namespace dpp {

int wgrad3_enabled() {      // read at every dpp_wgrad_group_create (tests switch it): bit 0 = 3x3 kernel, bit 1 = 1x1 kernel
    const char *e = getenv("DPP_WGRAD_SIMT3");
    return e ? atoi(e) : 3;
}
bool wgrad3_supported(const dpp_conv_desc *d);

// which layers these kernels take out of the tensor-core groups: 1 = 3x3 (k_wgrad3), 2 = 1x1 (k_wgrad1), 0 = none
int wgrad3_kind(const dpp_conv_desc *d) {
    if (!wgrad3_enabled()) return 0;
    if (d->k == 1 && d->stride == 1 && d->pad == 0 && ((d->Cin == 64 && d->Cout == 16) || (d->Cin == 16 && d->Cout == 64)) &&
        (wgrad3_enabled() & 2))
        return 2;
    return wgrad3_supported(d) ? 1 : 0;
}

bool wgrad3_supported(const dpp_conv_desc *d) {
    if (!(wgrad3_enabled() & 1)) return false;
    if (d->k != 3 || d->stride != 1 || d->pad != 1 || d->Cin != d->Cout || (d->Cin != 16 && d->Cin != 32)) return false;
    if (d->Ho != d->H || d->Wo != d->W) return false;
    const int bf = ((R3 + 2) * (d->W + 2) + R3 * d->W) * d->Cin;
    return 2 * bf * (int)sizeof(float) <= 200 * 1024;
}

int wgrad3_create(const std::vector<W3Layer> &layers, void **handle_out) {
    W3Group *g = new W3Group();
    memset(&g->args16, 0, sizeof(W3Args)); memset(&g->args32, 0, sizeof(W3Args));
    memset(&g->args1a, 0, sizeof(W3Args)); memset(&g->args1b, 0, sizeof(W3Args));
    if (build<16>(layers, g->args16, g->grid16, g->smem16, g->allocs) != 0 || build<32>(layers, g->args32, g->grid32, g->smem32, g->allocs) != 0 ||
        build1<64, 16>(layers, g->args1a, g->grid1a, g->smem1a, g->allocs) != 0 ||
        build1<16, 64>(layers, g->args1b, g->grid1b, g->smem1b, g->allocs) != 0) {
        for (void *p : g->allocs) cudaFree(p);
        delete g;
        return -1;
    }
    *handle_out = g;
    return 0;
}

int wgrad3_launches(void *handle) {
    W3Group *g = reinterpret_cast<W3Group *>(handle);
    return (g->grid16 > 0) + (g->grid32 > 0) + (g->grid1a > 0) + (g->grid1b > 0);
}

int wgrad3_run(void *handle, cudaStream_t st) {
    W3Group *g = reinterpret_cast<W3Group *>(handle);
    if (g->grid16 > 0) k_wgrad3<16><<<g->grid16, W3_THREADS, g->smem16, st>>>(g->args16);
    if (g->grid32 > 0) k_wgrad3<32><<<g->grid32, W3_THREADS, g->smem32, st>>>(g->args32);
    if (g->grid1a > 0) k_wgrad1<64, 16><<<g->grid1a, W3_THREADS, g->smem1a, st>>>(g->args1a);
    if (g->grid1b > 0) k_wgrad1<16, 64><<<g->grid1b, W3_THREADS, g->smem1b, st>>>(g->args1b);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

void wgrad3_destroy(void *handle) {
    W3Group *g = reinterpret_cast<W3Group *>(handle);
    for (void *p : g->allocs) cudaFree(p);
    delete g;
}

}  // namespace dpp
