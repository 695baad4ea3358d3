// HiddenLayer (net/hiddenlayer.py:136-154): y = act(x W + b), and its backward.  fp32 SIMT GEMM
// (precision 0).  The big one is FC0 (16384 -> 1024, net/resnet.py:141-145): at batch 128 every
// one of its three GEMMs is bound by streaming the 67 MB weight (or weight-gradient) matrix
// through HBM once, so the kernels below are organised around coalesced, split-K streaming of W.
#include "common.cuh"

using namespace dpp;

int dpp_gemm_tc(const float *A, const float *B, float *C, int M, int N, int K, int64_t lda, int64_t ldb, int a_trans,
                int b_trans, int precision, void *stream);
namespace dpp {
// fc_stream.cu: TMA-fed streaming 3xTF32 GEMM, C[n*ldc + m] (+)= sum_k A(m,k) B(n,k); DPP_ENOTSUP when the layout rules it out
struct FcEpilogue { const float *bias; const float *mask; float scale; int relu; };
int fc_stream_gemm(const float *A, int a_mn, int64_t lda, const float *B, int b_mn, int64_t ldb, float *C, int64_t ldc, int M,
                   int N, int K, int add, const FcEpilogue *epi, int *epi_done, cudaStream_t st);
}

namespace {

// C[M][N] (+)= sum_k A(m,k) B(k,n);  A(m,k) = A[m*am + k*ak], B(k,n) = B[k*bk + n*bn].
// 64x64x16 tiles, 256 threads, 4x4 micro-tiles.  gridDim.z = split-K (atomicAdd into C).
struct GArgs {
    const float *A; const float *B; float *C;
    int M, N, K;
    int64_t am, ak, bk, bn;
    int kchunk;       // K range per z-slice
    int atomic;       // 1: atomicAdd, 0: C += acc (single writer)
};

__global__ void __launch_bounds__(256)
k_gemm(GArgs g) {
    __shared__ __align__(16) float As[16][64 + 4];
    __shared__ __align__(16) float Bs[16][64 + 4];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int kb = blockIdx.z * g.kchunk;
    int ke = kb + g.kchunk; if (ke > g.K) ke = g.K;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // loader index maps: walk the contiguous dimension with consecutive threads
    const bool a_kcontig = (g.ak == 1);
    const bool b_ncontig = (g.bn == 1);
    for (int k0 = kb; k0 < ke; k0 += 16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int i = tid + j * 256;
            int ml, kl;
            if (a_kcontig) { kl = i % 16; ml = i / 16; } else { ml = i % 64; kl = i / 64; }
            int m = m0 + ml, k = k0 + kl;
            As[kl][ml] = (m < g.M && k < ke) ? g.A[m * g.am + k * g.ak] : 0.f;
            int nl, kl2;
            if (b_ncontig) { nl = i % 64; kl2 = i / 64; } else { kl2 = i % 16; nl = i / 16; }
            int n = n0 + nl, k2 = k0 + kl2;
            Bs[kl2][nl] = (n < g.N && k2 < ke) ? g.B[k2 * g.bk + n * g.bn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
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
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float *c = g.C + (size_t)m * g.N + n;
            if (g.atomic) atomicAdd(c, acc[i][j]); else *c += acc[i][j];
        }
    }
}

int run_gemm(const float *A, const float *B, float *C, int M, int N, int K, int64_t am, int64_t ak, int64_t bk,
             int64_t bn, cudaStream_t st, int precision = 0) {
    if (precision != 0) {
        const int a_trans = (ak != 1), b_trans = (bn == 1);
        int rc = dpp_gemm_tc(A, B, C, M, N, K, a_trans ? ak : am, b_trans ? bk : bn, a_trans, b_trans, precision, st);
        if (rc != DPP_ENOTSUP) return rc;
    }
    GArgs g{A, B, C, M, N, K, am, ak, bk, bn, K, 0};
    int tiles = cdiv(M, 64) * cdiv(N, 64);
    int split = 1;
    if (tiles < 148 * 2) {
        split = (148 * 3 + tiles - 1) / tiles;
        int maxs = K / 64; if (maxs < 1) maxs = 1;
        if (split > maxs) split = maxs;
    }
    int kchunk = ((K + split - 1) / split + 15) / 16 * 16;
    split = (K + kchunk - 1) / kchunk;
    g.kchunk = kchunk; g.atomic = split > 1;
    dim3 grid(cdiv(N, 64), cdiv(M, 64), split);
    k_gemm<<<grid, 256, 0, st>>>(g);
    return 0;
}

__global__ void k_fc_epilogue(float *__restrict__ y, const float *__restrict__ bias, const float *__restrict__ mask,
                              float scale, int relu, int64_t total, int n_out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float v = y[i] + bias[i % n_out];
        if (relu) v = fmaxf(v, 0.f);
        if (mask) v *= mask[i];
        y[i] = v * scale;
    }
}

// dpre = dy * mask * scale * [y > 0]; db[n] += sum_b dpre
// block = 32 output columns x 16 row groups (coalesced 128-byte rows); the row groups meet in shared memory
__global__ void __launch_bounds__(512)
k_fc_bwd_pre(const float *__restrict__ y, const float *__restrict__ dy, const float *__restrict__ mask,
             float scale, int relu, float *__restrict__ dpre, float *__restrict__ db, int B, int n_out) {
    __shared__ float s_part[16][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + tx;
    float s = 0.f;
    if (n < n_out)
        for (int b = ty; b < B; b += 16) {
            size_t i = (size_t)b * n_out + n;
            float g = dy[i] * scale;
            if (mask) g *= mask[i];
            if (relu && !(y[i] > 0.f)) g = 0.f;
            dpre[i] = g;
            s += g;
        }
    s_part[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && n < n_out) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) t += s_part[j][tx];
        db[n] += t;
    }
}

}  // namespace

extern "C" int dpp_fc_fwd(const float *x, const float *w, const float *bias, float *y, int B, int n_in, int n_out,
                          int relu, const float *mask, float scale_out, int precision, void *stream) {
    DPP_CHECK_ARG(x && w && bias && y && B > 0 && n_in > 0 && n_out > 0);
    // y[b][o] = sum_i W[i][o] x[b][i]: lanes = output units, columns = samples
    const FcEpilogue epi{bias, mask, scale_out, relu};
    int epi_done = 0;
    int rc = precision == 1 ? fc_stream_gemm(w, 1, n_out, x, 0, n_in, y, n_out, n_out, B, n_in, 0, &epi, &epi_done, S(stream)) : DPP_ENOTSUP;
    if (rc == DPP_ENOTSUP) {
        DPP_CUDA(cudaMemsetAsync(y, 0, sizeof(float) * (size_t)B * n_out, S(stream)));
        run_gemm(x, w, y, B, n_out, n_in, n_in, 1, n_out, 1, S(stream), precision);
    } else if (rc != DPP_OK) {
        return dpp::fail(rc, "%s: streaming GEMM launch failed", __func__);
    }
    DPP_LAUNCH_CHECK();
    if (!epi_done) {        // (the streaming GEMM's split reduction applies bias / activation / mask itself)
        int64_t total = (int64_t)B * n_out;
        int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
        k_fc_epilogue<<<blocks, 256, 0, S(stream)>>>(y, bias, mask, scale_out, relu, total, n_out);
        DPP_LAUNCH_CHECK();
    }
    return DPP_OK;
}

extern "C" int dpp_fc_bwd_ex(const float *x, const float *w, const float *y, const float *dy, float *dw, float *db,
                             float *dx, float *scratch, int B, int n_in, int n_out, int relu, const float *mask,
                             float scale_out, int precision, int flags, void *stream) {
    DPP_CHECK_ARG(x && w && y && dy && dw && db && scratch && B > 0);
    const int assign = (flags & DPP_FC_DW_ASSIGN) ? 1 : 0;
    k_fc_bwd_pre<<<cdiv(n_out, 32), 512, 0, S(stream)>>>(y, dy, mask, scale_out, relu, scratch, db, B, n_out);
    DPP_LAUNCH_CHECK();
    // dW[i][o] (+)= sum_b x[b][i] dpre[b][o]: lanes = output units, columns = input units, reduction over the samples
    int rc = precision == 1 ? fc_stream_gemm(scratch, 1, n_out, x, 1, n_in, dw, n_out, n_out, n_in, B, !assign, nullptr, nullptr, S(stream))
                            : DPP_ENOTSUP;
    if (rc == DPP_ENOTSUP) {
        if (assign) DPP_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)n_in * n_out, S(stream)));
        run_gemm(x, scratch, dw, n_in, n_out, B, 1, n_in, n_out, 1, S(stream), precision);
    } else if (rc != DPP_OK) {
        return dpp::fail(rc, "%s: streaming GEMM (dW) launch failed", __func__);
    }
    DPP_LAUNCH_CHECK();
    if (dx) {
        // dx[b][i] = sum_o W[i][o] dpre[b][o]: lanes = input units, columns = samples
        rc = precision == 1 ? fc_stream_gemm(w, 0, n_out, scratch, 0, n_out, dx, n_in, n_in, B, n_out, 0, nullptr, nullptr, S(stream))
                            : DPP_ENOTSUP;
        if (rc == DPP_ENOTSUP) {
            DPP_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)B * n_in, S(stream)));
            run_gemm(scratch, w, dx, B, n_in, n_out, n_out, 1, 1, n_out, S(stream), precision);
        } else if (rc != DPP_OK) {
            return dpp::fail(rc, "%s: streaming GEMM (dx) launch failed", __func__);
        }
        DPP_LAUNCH_CHECK();
    }
    return DPP_OK;
}

extern "C" int dpp_fc_bwd(const float *x, const float *w, const float *y, const float *dy, float *dw, float *db,
                          float *dx, float *scratch, int B, int n_in, int n_out, int relu, const float *mask,
                          float scale_out, int precision, void *stream) {
    return dpp_fc_bwd_ex(x, w, y, dy, dw, db, dx, scratch, B, n_in, n_out, relu, mask, scale_out, precision, 0, stream);
}
