// fp32 SIMT implicit-GEMM convolution: forward, backward-data, backward-weights, with the
// BatchNorm/ReLU layers that surround every ConvLayer of the ResNet fused in.
//
// Reference call sites: net/convlayer.py:230-235 (theano conv2d, border 'half', stride 1|2),
// net/batchnormlayer.py:154-192 + net/nonlinearitylayer.py:119 (prologue), the residual sums of
// net/resnet.py:379,414 (epilogue) and T.grad of all of it (trainer/poseregnettrainer.py:111).
//
// This is precision mode 0 ("exact fp32"): it is the parity anchor against the oracle and the
// fallback for shapes the tcgen05 path (conv_tc.cu) does not take.  GEMM view:
//   out[m][n] = sum_{r,s,c} T(in)[pix(m)*in_stride - pad + (r,s)][c] * B[(r,s,c)][n]
// forward : in = x, T = BN+ReLU prologue, B = W_kc, epilogue bias (+residual) (+BN sums)
// dgrad   : in = dy, B[(r',s',o)][c] = W_kc[flip(r',s'), c][o], epilogue ReLU mask + BN-bwd sums
// Tiles: 64 pixels x BN channels x 16 k, 4x4 register micro-tiles; persistent CTAs loop over
// pixel tiles so the per-channel fp64 statistics are reduced once per CTA.
#include "common.cuh"

using namespace dpp;

namespace {

constexpr int BM = 64;
constexpr int BK = 16;

struct IGArgs {
    const float *in;      // [N, Hin, Win, Cin]
    const float *w;       // KC weights of the layer
    float *out;           // [N, Hout, Wout, Cn]
    int N, Hin, Win, Cin; // gathered tensor
    int Hg, Wg;           // GEMM pixel grid (rows = N*Hg*Wg)
    int Hout, Wout, Cn;   // output tensor / GEMM N
    int k, pad, in_stride, out_stride;
    int wmode;            // 0: B = w[(r,s,c)][n] ; 1: dgrad B[(r,s,o)][c] = w[flip(r,s), c][o]
    int wCin, wCout;      // the layer's own Cin/Cout (for wmode 1 indexing)
    // forward prologue / epilogue
    dpp_bn_ref in_bn; int has_in_bn;
    const float *bias; const float *residual; double *out_stats;
    // dgrad epilogue
    int accumulate; dpp_bn_ref mask_bn; int has_mask; const float *x_pre; double *dz_stats;
};

template <int BN_>
__global__ void __launch_bounds__((BM / 4) * (BN_ / 4))
k_igemm(IGArgs a) {
    constexpr int T = (BM / 4) * (BN_ / 4);
    constexpr int TX = BN_ / 4;                 // threads along n
    constexpr int A_F4 = BM * BK / 4;           // float4 per A tile (256)
    constexpr int A_PER_T = (A_F4 + T - 1) / T; // float4 loads per thread
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN_];
    __shared__ float s_scale[256], s_shift[256];       // input-BN prologue (Cin <= 256)
    __shared__ float s_mscale[BN_], s_mshift[BN_], s_mmean[BN_], s_mistd[BN_];  // dgrad mask BN (this n-tile)
    __shared__ float s_red[2][BM / 4][BN_];
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int n0 = blockIdx.y * BN_;
    const int M = a.N * a.Hg * a.Wg;
    const int mtiles = (M + BM - 1) / BM;

    if (a.has_in_bn)
        for (int c = tid; c < a.Cin; c += T) bn_scale_shift(a.in_bn, c, a.Cin, s_scale[c], s_shift[c]);
    if (a.has_mask)
        for (int c = tid; c < BN_; c += T) {
            float mean, istd;
            bn_mean_istd(a.mask_bn, n0 + c, a.Cn, mean, istd);
            float sc = a.mask_bn.gamma[n0 + c] * istd;
            s_mscale[c] = sc; s_mshift[c] = a.mask_bn.beta[n0 + c] - mean * sc;
            s_mmean[c] = mean; s_mistd[c] = istd;
        }
    // per-CTA fp64 column statistics, owned by threads tid < BN_
    double acc_s0 = 0.0, acc_s1 = 0.0;
    const bool want_stats = (a.out_stats != nullptr) || (a.dz_stats != nullptr);
    __syncthreads();

    for (int mt = blockIdx.x; mt < mtiles; mt += gridDim.x) {
        const int m0 = mt * BM;
        // rows this thread loads: row(q) = q / 4, q = tid + j*T
        int rn[A_PER_T], rh[A_PER_T], rw[A_PER_T];
        bool rv[A_PER_T];
#pragma unroll
        for (int j = 0; j < A_PER_T; ++j) {
            int q = tid + j * T;
            int m = m0 + q / 4;
            rv[j] = (q < A_F4) && (m < M);
            int mm = rv[j] ? m : 0;
            rw[j] = mm % a.Wg; rh[j] = (mm / a.Wg) % a.Hg; rn[j] = mm / (a.Wg * a.Hg);
        }
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

        for (int r = 0; r < a.k; ++r)
            for (int s = 0; s < a.k; ++s) {
                const float *rp[A_PER_T];
                bool ok[A_PER_T];
#pragma unroll
                for (int j = 0; j < A_PER_T; ++j) {
                    int hi = rh[j] * a.in_stride - a.pad + r, wi = rw[j] * a.in_stride - a.pad + s;
                    ok[j] = rv[j] && hi >= 0 && hi < a.Hin && wi >= 0 && wi < a.Win;
                    rp[j] = a.in + (((size_t)rn[j] * a.Hin + (ok[j] ? hi : 0)) * a.Win + (ok[j] ? wi : 0)) * a.Cin;
                }
                const int tap = r * a.k + s;
                for (int c0 = 0; c0 < a.Cin; c0 += BK) {
                    // ---- A tile: gather + BN/ReLU prologue, stored k-major
#pragma unroll
                    for (int j = 0; j < A_PER_T; ++j) {
                        int q = tid + j * T;
                        if (q < A_F4) {
                            int ml = q / 4, kq = (q % 4) * 4;
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (ok[j]) {
                                v = *reinterpret_cast<const float4 *>(rp[j] + c0 + kq);
                                if (a.has_in_bn) {
                                    int c = c0 + kq;
                                    v.x = fmaf(v.x, s_scale[c], s_shift[c]);
                                    v.y = fmaf(v.y, s_scale[c + 1], s_shift[c + 1]);
                                    v.z = fmaf(v.z, s_scale[c + 2], s_shift[c + 2]);
                                    v.w = fmaf(v.w, s_scale[c + 3], s_shift[c + 3]);
                                    if (a.in_bn.relu) {
                                        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f);
                                        v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                                    }
                                }
                            }
                            As[kq][ml] = v.x; As[kq + 1][ml] = v.y; As[kq + 2][ml] = v.z; As[kq + 3][ml] = v.w;
                        }
                    }
                    // ---- B tile
                    for (int i = tid; i < BK * BN_; i += T) {
                        int kk = i / BN_, nn = i % BN_;
                        float wv;
                        if (a.wmode == 0) {
                            wv = a.w[(size_t)(tap * a.Cin + c0 + kk) * a.Cn + n0 + nn];
                        } else {
                            // gathered tensor is dy: its channel index is the layer's o; n is the layer's c
                            int ftap = (a.k - 1 - r) * a.k + (a.k - 1 - s);
                            wv = a.w[(size_t)(ftap * a.wCin + n0 + nn) * a.wCout + c0 + kk];
                        }
                        Bs[kk][nn] = wv;
                    }
                    __syncthreads();
#pragma unroll
                    for (int kk = 0; kk < BK; ++kk) {
                        float4 av = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
                        float4 bv = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
                        float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
                    }
                    __syncthreads();
                }
            }

        // ---- epilogue
        float cs0[4] = {0.f, 0.f, 0.f, 0.f}, cs1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m = m0 + ty * 4 + i;
            if (m >= M) continue;
            int wo = m % a.Wg, ho = (m / a.Wg) % a.Hg, n = m / (a.Wg * a.Hg);
            size_t ob = (((size_t)n * a.Hout + ho * a.out_stride) * a.Wout + wo * a.out_stride) * a.Cn + n0 + tx * 4;
            float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
            if (a.wmode == 0) {
                if (a.bias) {
                    float4 b = *reinterpret_cast<const float4 *>(a.bias + n0 + tx * 4);
                    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
                }
                if (a.residual) {
                    float4 rr = *reinterpret_cast<const float4 *>(a.residual + ob);
                    v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
                }
                if (a.out_stats) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { cs0[j] += v[j]; cs1[j] += v[j] * v[j]; }
                }
            } else {
                if (a.accumulate) {
                    float4 e = *reinterpret_cast<const float4 *>(a.out + ob);
                    v[0] += e.x; v[1] += e.y; v[2] += e.z; v[3] += e.w;
                }
                if (a.has_mask) {
                    float4 xp = *reinterpret_cast<const float4 *>(a.x_pre + ob);
                    float xr[4] = {xp.x, xp.y, xp.z, xp.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        int cl = tx * 4 + j;
                        float pre = fmaf(xr[j], s_mscale[cl], s_mshift[cl]);
                        float dz = (pre > 0.f) ? v[j] : 0.f;
                        float xh = (xr[j] - s_mmean[cl]) * s_mistd[cl];
                        v[j] = dz;
                        cs0[j] += dz; cs1[j] += dz * xh;
                    }
                }
            }
            *reinterpret_cast<float4 *>(a.out + ob) = make_float4(v[0], v[1], v[2], v[3]);
        }
        if (want_stats) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { s_red[0][ty][tx * 4 + j] = cs0[j]; s_red[1][ty][tx * 4 + j] = cs1[j]; }
            __syncthreads();
            if (tid < BN_) {
                float t0 = 0.f, t1 = 0.f;
#pragma unroll
                for (int i = 0; i < BM / 4; ++i) { t0 += s_red[0][i][tid]; t1 += s_red[1][i][tid]; }
                acc_s0 += (double)t0; acc_s1 += (double)t1;
            }
            __syncthreads();
        }
    }
    if (want_stats && tid < BN_) {
        double *st = a.out_stats ? a.out_stats : a.dz_stats;
        atomicAdd(&st[n0 + tid], acc_s0);
        atomicAdd(&st[a.Cn + n0 + tid], acc_s1);
    }
}

template <int BN_>
int launch_igemm(const IGArgs &a, cudaStream_t st) {
    constexpr int T = (BM / 4) * (BN_ / 4);
    static_assert(T >= BN_, "stat owners");
    int M = a.N * a.Hg * a.Wg;
    int mtiles = (M + BM - 1) / BM;
    int ntiles = a.Cn / BN_;
    int per_sm = 2048 / T; if (per_sm > 8) per_sm = 8;
    int gx = 148 * per_sm / ntiles; if (gx < 1) gx = 1; if (gx > mtiles) gx = mtiles;
    dim3 grid(gx, ntiles);
    k_igemm<BN_><<<grid, T, 0, st>>>(a);
    return 0;
}

int dispatch_igemm(const IGArgs &a, cudaStream_t st) {
    if (a.Cn % 64 == 0) return launch_igemm<64>(a, st);
    if (a.Cn % 32 == 0) return launch_igemm<32>(a, st);
    if (a.Cn % 16 == 0) return launch_igemm<16>(a, st);
    return -1;
}

// ---------------------------------------------------------------------------------------
// wgrad: dW[(r,s,c)][o] += sum_p a[p*stride - pad + (r,s)][c] * dy[p][o],  db[o] += sum_p dy[p][o]
// 64 (kdim) x 64 (cout) tile per CTA, reduction over a slice of the pixels, fp32 atomics out.
// ---------------------------------------------------------------------------------------
struct WGArgs {
    const float *x; const float *dy; float *dw; float *db;
    int N, H, W, Cin, Cout, k, stride, pad, Ho, Wo;
    dpp_bn_ref in_bn; int has_in_bn;
    int cotiles; int pix_per_cta;
};

__global__ void __launch_bounds__(256)
k_wgrad(WGArgs a) {
    constexpr int TK = 64, TN = 64, BP = 16;
    __shared__ __align__(16) float As[BP][TK];
    __shared__ __align__(16) float Bs[BP][TN];
    __shared__ float s_scale[256], s_shift[256];
    __shared__ float s_db[16][TN];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int kt = blockIdx.y / a.cotiles, ct = blockIdx.y % a.cotiles;
    const int kd0 = kt * TK, o0 = ct * TN;
    const int Kw = a.k * a.k * a.Cin;
    const int P = a.N * a.Ho * a.Wo;
    if (a.has_in_bn)
        for (int c = tid; c < a.Cin; c += 256) bn_scale_shift(a.in_bn, c, a.Cin, s_scale[c], s_shift[c]);
    __syncthreads();

    // this thread's slice of the A tile: pixel pl, kdim quad kq
    const int pl = tid / 16, kq = (tid % 16) * 4;
    const int kd = kd0 + kq;
    const bool kvalid = kd < Kw;
    const int tap = kvalid ? kd / a.Cin : 0, c = kvalid ? kd % a.Cin : 0;
    const int r = tap / a.k, s = tap % a.k;
    const int oq = (tid % 16) * 4;            // B tile column quad
    const bool ovalid = (o0 + oq) < a.Cout;
    float dbp[4] = {0.f, 0.f, 0.f, 0.f};

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int p_begin = blockIdx.x * a.pix_per_cta;
    int p_end = p_begin + a.pix_per_cta; if (p_end > P) p_end = P;
    for (int p0 = p_begin; p0 < p_end; p0 += BP) {
        int p = p0 + pl;
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < p_end) {
            int wo = p % a.Wo, ho = (p / a.Wo) % a.Ho, n = p / (a.Wo * a.Ho);
            if (kvalid) {
                int hi = ho * a.stride - a.pad + r, wi = wo * a.stride - a.pad + s;
                if (hi >= 0 && hi < a.H && wi >= 0 && wi < a.W) {
                    av = *reinterpret_cast<const float4 *>(a.x + (((size_t)n * a.H + hi) * a.W + wi) * a.Cin + c);
                    if (a.has_in_bn) {
                        av.x = fmaf(av.x, s_scale[c], s_shift[c]);
                        av.y = fmaf(av.y, s_scale[c + 1], s_shift[c + 1]);
                        av.z = fmaf(av.z, s_scale[c + 2], s_shift[c + 2]);
                        av.w = fmaf(av.w, s_scale[c + 3], s_shift[c + 3]);
                        if (a.in_bn.relu) {
                            av.x = fmaxf(av.x, 0.f); av.y = fmaxf(av.y, 0.f);
                            av.z = fmaxf(av.z, 0.f); av.w = fmaxf(av.w, 0.f);
                        }
                    }
                }
            }
            if (ovalid) bv = *reinterpret_cast<const float4 *>(a.dy + (size_t)p * a.Cout + o0 + oq);
        }
        *reinterpret_cast<float4 *>(&As[pl][kq]) = av;
        *reinterpret_cast<float4 *>(&Bs[pl][oq]) = bv;
        dbp[0] += bv.x; dbp[1] += bv.y; dbp[2] += bv.z; dbp[3] += bv.w;
        __syncthreads();
#pragma unroll
        for (int pp = 0; pp < BP; ++pp) {
            float4 a4 = *reinterpret_cast<const float4 *>(&As[pp][ty * 4]);
            float4 b4 = *reinterpret_cast<const float4 *>(&Bs[pp][tx * 4]);
            float ar[4] = {a4.x, a4.y, a4.z, a4.w}, br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int kdd = kd0 + ty * 4 + i;
        if (kdd >= Kw) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int o = o0 + tx * 4 + j;
            if (o < a.Cout) atomicAdd(&a.dw[(size_t)kdd * a.Cout + o], acc[i][j]);
        }
    }
    if (kt == 0 && a.db) {
#pragma unroll
        for (int j = 0; j < 4; ++j) s_db[pl][oq + j] = dbp[j];
        __syncthreads();
        if (tid < TN && o0 + tid < a.Cout) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) t += s_db[i][tid];
            atomicAdd(&a.db[o0 + tid], t);
        }
    }
}

int check_desc(const dpp_conv_desc *d) {
    if (!d) return -1;
    if (d->N <= 0 || d->H <= 0 || d->W <= 0) return -1;
    if (d->Cin % 16 || d->Cout % 16 || d->Cin > 256 || d->Cout > 256) return -1;
    if (!(d->k == 1 || d->k == 3 || d->k == 5)) return -1;
    if (!(d->stride == 1 || (d->stride == 2 && d->k == 1))) return -1;
    if (d->pad != d->k / 2) return -1;
    int Ho = (d->H + 2 * d->pad - d->k) / d->stride + 1, Wo = (d->W + 2 * d->pad - d->k) / d->stride + 1;
    if (Ho != d->Ho || Wo != d->Wo) return -1;
    return 0;
}

}  // namespace

// tcgen05 path (conv_tc.cu); returns DPP_ENOTSUP when it does not take the shape
int dpp_conv2d_fwd_tc(const dpp_conv_desc *d, const float *x, const dpp_bn_ref *in_bn, const float *w,
                      const float *bias, const float *residual, float *y, double *out_stats, void *stream);
int dpp_conv2d_dgrad_tc(const dpp_conv_desc *d, const float *dy, float *dx, int accumulate, const dpp_bn_ref *mask_bn,
                        const float *x_pre, double *dz_stats, void *stream);
int dpp_conv2d_wgrad_tc(const dpp_conv_desc *d, const float *x, const dpp_bn_ref *in_bn, const float *dy, float *dw,
                        float *db, void *stream);

extern "C" int dpp_conv2d_fwd(const dpp_conv_desc *d, const float *x, const dpp_bn_ref *in_bn, const float *w,
                              const float *bias, const float *residual, float *y, double *out_stats,
                              void *stream) {
    DPP_CHECK_ARG(check_desc(d) == 0 && x && w && y);
    if (d->precision != 0) {
        int rc = dpp_conv2d_fwd_tc(d, x, in_bn, w, bias, residual, y, out_stats, stream);
        if (rc != DPP_ENOTSUP) return rc;
    }
    IGArgs a;
    memset(&a, 0, sizeof(a));
    a.in = x; a.w = w; a.out = y;
    a.N = d->N; a.Hin = d->H; a.Win = d->W; a.Cin = d->Cin;
    a.Hg = d->Ho; a.Wg = d->Wo; a.Hout = d->Ho; a.Wout = d->Wo; a.Cn = d->Cout;
    a.k = d->k; a.pad = d->pad; a.in_stride = d->stride; a.out_stride = 1;
    a.wmode = 0; a.wCin = d->Cin; a.wCout = d->Cout;
    if (in_bn) { a.in_bn = *in_bn; a.has_in_bn = 1; }
    a.bias = bias; a.residual = residual; a.out_stats = out_stats;
    DPP_CHECK_ARG(dispatch_igemm(a, S(stream)) == 0);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

extern "C" int dpp_conv2d_dgrad(const dpp_conv_desc *d, const float *dy, const float *w, float *dx, int accumulate,
                                const dpp_bn_ref *mask_bn, const float *x_pre, double *dz_stats, void *stream) {
    DPP_CHECK_ARG(check_desc(d) == 0 && dy && w && dx);
    DPP_CHECK_ARG(!mask_bn || (x_pre && dz_stats));
    if (d->precision != 0) {
        int rc = dpp_conv2d_dgrad_tc(d, dy, dx, accumulate, mask_bn, x_pre, dz_stats, stream);
        if (rc != DPP_ENOTSUP) return rc;
    }
    IGArgs a;
    memset(&a, 0, sizeof(a));
    a.in = dy; a.w = w; a.out = dx;
    a.N = d->N; a.Hin = d->Ho; a.Win = d->Wo; a.Cin = d->Cout;      // gathered tensor = dy
    a.Hout = d->H; a.Wout = d->W; a.Cn = d->Cin;
    a.k = d->k; a.pad = d->k - 1 - d->pad; a.in_stride = 1;
    if (d->stride == 1) { a.Hg = d->H; a.Wg = d->W; a.out_stride = 1; }
    else { a.Hg = d->Ho; a.Wg = d->Wo; a.out_stride = d->stride; }  // 1x1/s2: scatter to even positions
    a.wmode = 1; a.wCin = d->Cin; a.wCout = d->Cout;
    a.accumulate = accumulate;
    if (mask_bn) { a.mask_bn = *mask_bn; a.has_mask = 1; a.x_pre = x_pre; a.dz_stats = dz_stats; }
    DPP_CHECK_ARG(dispatch_igemm(a, S(stream)) == 0);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

extern "C" int dpp_conv2d_wgrad(const dpp_conv_desc *d, const float *x, const dpp_bn_ref *in_bn, const float *dy,
                                float *dw, float *db, void *stream) {
    DPP_CHECK_ARG(check_desc(d) == 0 && x && dy && dw);
    if (d->precision != 0) {
        int rc = dpp_conv2d_wgrad_tc(d, x, in_bn, dy, dw, db, stream);
        if (rc != DPP_ENOTSUP) return rc;
    }
    WGArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.dy = dy; a.dw = dw; a.db = db;
    a.N = d->N; a.H = d->H; a.W = d->W; a.Cin = d->Cin; a.Cout = d->Cout;
    a.k = d->k; a.stride = d->stride; a.pad = d->pad; a.Ho = d->Ho; a.Wo = d->Wo;
    if (in_bn) { a.in_bn = *in_bn; a.has_in_bn = 1; }
    int Kw = d->k * d->k * d->Cin;
    int ktiles = (Kw + 63) / 64;
    a.cotiles = (d->Cout + 63) / 64;
    int P = d->N * d->Ho * d->Wo;
    int ytiles = ktiles * a.cotiles;
    int want = 148 * 4 / ytiles; if (want < 1) want = 1;
    int ppc = (P + want - 1) / want;
    ppc = ((ppc + 15) / 16) * 16; if (ppc < 64) ppc = 64;
    a.pix_per_cta = ppc;
    dim3 grid((P + ppc - 1) / ppc, ytiles);
    k_wgrad<<<grid, 256, 0, S(stream)>>>(a);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}
