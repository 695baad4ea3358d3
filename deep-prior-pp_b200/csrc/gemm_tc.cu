// tcgen05 GEMM for the HiddenLayers (reference net/hiddenlayer.py:136-154 and its T.grad):
//   D[M][N] += sum_k A(m,k) * B(n,k)       (3xTF32 or TF32, fp32 accumulation in TMEM)
// with each operand either k-contiguous in memory (rows copied straight into the K-major
// SWIZZLE_128B tile) or m/n-contiguous (transposed on the fly, 4-byte scatter, as in k_wgrad_tc):
//   forward  y  = x W        : A = x  [B][n_in]   straight,   B(n,k) = W[k][n]   transposed
//   backward dx = dpre W^T   : A = dpre [B][n_out] straight,  B(n,k) = W[n][k]   straight
//   backward dW = x^T dpre   : A(m,k) = x[k][m]   transposed, B(n,k) = dpre[k][n] transposed
// FC0 (16384 x 1024) is a 67 MB weight stream per GEMM at batch 128: split-K keeps >= 128 CTAs busy
// and the epilogue adds partial tiles with red.global.add.f32 (outputs are pre-zeroed by the caller).
#include "tc_common.cuh"

using namespace dpp;
using namespace dpp::tc;

namespace {

constexpr int TM = 128;
constexpr int KC = 32;
constexpr int NSTAGE = 3;
constexpr int NTHREADS = 288;

struct GTArgs {
    const float *A; const float *B; float *C;
    int M, N, K;
    int64_t lda, ldb;      // leading dimension (elements) of the memory-contiguous direction's rows
    int ntiles, splits, chunks_per_split;
};

// store a 16-byte piece `v` (4 consecutive k) of tile row `row`, chunk `ch` (0..7)
template <int ROWS, int PASSES>
__device__ __forceinline__ void st_piece(unsigned char *tile, int row, int ch, float4 v) {
    const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) << 4);
    uint4 h;
    h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
    *reinterpret_cast<uint4 *>(tile + off) = h;
    if (PASSES > 1) {
        uint4 l;
        l.x = lo_tf32(v.x, h.x); l.y = lo_tf32(v.y, h.y);
        l.z = lo_tf32(v.z, h.z); l.w = lo_tf32(v.w, h.w);
        *reinterpret_cast<uint4 *>(tile + ROWS * 128 + off) = l;
    }
}
// store one element (row, k-column j)
template <int ROWS, int PASSES>
__device__ __forceinline__ void st_elem(unsigned char *tile, int row, int j, float e) {
    const int off = (row >> 3) * 1024 + (row & 7) * 128 + (((j >> 2) ^ (row & 7)) << 4) + ((j & 3) << 2);
    const uint32_t h = to_tf32(e);
    *reinterpret_cast<uint32_t *>(tile + off) = h;
    if (PASSES > 1) *reinterpret_cast<uint32_t *>(tile + ROWS * 128 + off) = lo_tf32(e, h);
}

template <int BN, int PASSES, bool ATRANS, bool BTRANS>
__global__ void __launch_bounds__(NTHREADS, 1)
k_gemm_tc(GTArgs a) {
    constexpr int A_BYTES = PASSES * TM * 128, B_BYTES = PASSES * BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int BAR_OFF = NSTAGE * STAGE_BYTES;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still in the shared window
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto bar = [&](int i) { return sbase + BAR_OFF + 8 * i; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + BAR_OFF + 128);
    constexpr uint32_t TCOLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : 128);

    const int mtiles = (a.M + TM - 1) / TM;
    const int tile = blockIdx.x % (mtiles * a.ntiles), split = blockIdx.x / (mtiles * a.ntiles);
    const int m0 = (tile / a.ntiles) * TM, n0 = (tile % a.ntiles) * BN;
    const int total_chunks = (a.K + KC - 1) / KC;
    const int c_begin = split * a.chunks_per_split;
    int c_end = c_begin + a.chunks_per_split; if (c_end > total_chunks) c_end = total_chunks;
    const int nchunks = c_end > c_begin ? c_end - c_begin : 0;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar(s), 128); mbar_init(bar(NSTAGE + s), 1); }
        mbar_init(bar(2 * NSTAGE), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        uint32_t stage = 0, phase = 0;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ch = 0; ch < nchunks; ++ch) {
            const int k0 = (c_begin + ch) * KC;
            // ---- loads first (overlap the wait for the slot)
            float4 va[8];
            if (!ATRANS) {               // thread = row; 8 pieces along k
                const int m = m0 + tid;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const int k = k0 + g * 4;
                    va[g] = (m < a.M && k < a.K) ? *reinterpret_cast<const float4 *>(a.A + (size_t)m * a.lda + k) : z4;
                }
            } else {                     // thread = (k column j, row quarter q); 8 pieces along m
                const int j = tid & 31, q = tid >> 5, k = k0 + j;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const int m = m0 + q * 32 + g * 4;
                    va[g] = (k < a.K && m < a.M) ? *reinterpret_cast<const float4 *>(a.A + (size_t)k * a.lda + m) : z4;
                }
            }
            constexpr int BP = BN / 16;  // 16-byte pieces of B per thread
            float4 vb[BP];
            if (!BTRANS) {
#pragma unroll
                for (int g = 0; g < BP; ++g) {
                    const int idx = tid + g * 128, row = idx >> 3, chk = idx & 7;
                    const int n = n0 + row, k = k0 + chk * 4;
                    vb[g] = (n < a.N && k < a.K) ? *reinterpret_cast<const float4 *>(a.B + (size_t)n * a.ldb + k) : z4;
                }
            } else {
                const int j = tid & 31, q = tid >> 5, k = k0 + j;
#pragma unroll
                for (int g = 0; g < BP; ++g) {
                    const int n = n0 + q * (BN / 4) + g * 4;
                    vb[g] = (k < a.K && n < a.N) ? *reinterpret_cast<const float4 *>(a.B + (size_t)k * a.ldb + n) : z4;
                }
            }
            mbar_wait(bar(NSTAGE + stage), phase ^ 1);
            unsigned char *sA = smem + stage * STAGE_BYTES, *sB = sA + A_BYTES;
            if (!ATRANS) {
#pragma unroll
                for (int g = 0; g < 8; ++g) st_piece<TM, PASSES>(sA, tid, g, va[g]);
            } else {
                const int j = tid & 31, q = tid >> 5;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float e[4] = {va[g].x, va[g].y, va[g].z, va[g].w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) st_elem<TM, PASSES>(sA, q * 32 + g * 4 + t, j, e[t]);
                }
            }
            if (!BTRANS) {
#pragma unroll
                for (int g = 0; g < BP; ++g) {
                    const int idx = tid + g * 128;
                    st_piece<BN, PASSES>(sB, idx >> 3, idx & 7, vb[g]);
                }
            } else {
                const int j = tid & 31, q = tid >> 5;
#pragma unroll
                for (int g = 0; g < BP; ++g) {
                    const float e[4] = {vb[g].x, vb[g].y, vb[g].z, vb[g].w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) st_elem<BN, PASSES>(sB, q * (BN / 4) + g * 4 + t, j, e[t]);
                }
            }
            fence_proxy_async();
            mbar_arrive(bar(stage));
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 8) {
        if (nchunks > 0 && elect_one()) {      // ONE thread runs the role (see tc_common.cuh)
            constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            uint32_t stage = 0, phase = 0;
            for (int ch = 0; ch < nchunks; ++ch) {
                mbar_wait(bar(stage), phase);
                tc_fence_after();
                const uint32_t sa = sbase + stage * STAGE_BYTES;
                const uint64_t a0 = make_desc(sa), b0 = make_desc(sa + A_BYTES);
#pragma unroll
                for (int ks = 0; ks < KC / 8; ++ks) {
                    const uint64_t ah = a0 + 2 * ks, bh = b0 + 2 * ks;
                    const uint32_t first = (ch == 0 && ks == 0) ? 0u : 1u;
                    if (PASSES > 1) {
                        const uint64_t al = ah + ((TM * 128) >> 4), bl = bh + ((BN * 128) >> 4);
                        mma_tf32_ss_1t(tmem_base, ah, bl, IDESC, first);
                        mma_tf32_ss_1t(tmem_base, al, bh, IDESC, 1u);
                        mma_tf32_ss_1t(tmem_base, ah, bh, IDESC, 1u);
                    } else {
                        mma_tf32_ss_1t(tmem_base, ah, bh, IDESC, first);
                    }
                }
                mma_commit_1t(bar(NSTAGE + stage));
                if (ch == nchunks - 1) mma_commit_1t(bar(2 * NSTAGE));
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (nchunks > 0) {
        const int ew = warp - 4;
        const int m = m0 + ew * 32 + lane;
        mbar_wait(bar(2 * NSTAGE), 0);
        tc_fence_after();
#pragma unroll
        for (int cb = 0; cb < BN; cb += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + cb, v);
            if (m < a.M) {
                float *dst = a.C + (size_t)m * a.N + n0 + cb;
#pragma unroll
                for (int t = 0; t < 16; t += 4)
                    if (n0 + cb + t < a.N)      // N % 4 == 0: whole quads are in or out
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + t), "f"(v[t]), "f"(v[t + 1]),
                                     "f"(v[t + 2]), "f"(v[t + 3])
                                     : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
    }
}

template <int BN, int PASSES, bool AT, bool BT>
int launch_gemm_tc(GTArgs &a, cudaStream_t st) {
    constexpr int SMEM = NSTAGE * (PASSES * TM * 128 + PASSES * BN * 128) + 256 + 1024;
    static bool done = false;
    if (!done) {
        if (cudaFuncSetAttribute(k_gemm_tc<BN, PASSES, AT, BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess)
            return -1;
        done = true;
    }
    const int mtiles = (a.M + TM - 1) / TM;
    a.ntiles = (a.N + BN - 1) / BN;
    const int tiles = mtiles * a.ntiles;
    const int total_chunks = (a.K + KC - 1) / KC;
    int splits = (148 + tiles - 1) / tiles; if (splits < 1) splits = 1;
    if (splits > total_chunks / 4) splits = total_chunks / 4 > 0 ? total_chunks / 4 : 1;
    a.chunks_per_split = (total_chunks + splits - 1) / splits;
    a.splits = (total_chunks + a.chunks_per_split - 1) / a.chunks_per_split;
    k_gemm_tc<BN, PASSES, AT, BT><<<tiles * a.splits, NTHREADS, SMEM, st>>>(a);
    return 0;
}

template <bool AT, bool BT>
int dispatch(GTArgs &a, int precision, cudaStream_t st) {
    const bool p3 = precision == 1;
    if (a.N >= 128) return p3 ? launch_gemm_tc<128, 2, AT, BT>(a, st) : launch_gemm_tc<128, 1, AT, BT>(a, st);
    if (a.N > 32) return p3 ? launch_gemm_tc<64, 2, AT, BT>(a, st) : launch_gemm_tc<64, 1, AT, BT>(a, st);
    return p3 ? launch_gemm_tc<32, 2, AT, BT>(a, st) : launch_gemm_tc<32, 1, AT, BT>(a, st);
}

}  // namespace

// C[M][N] += A op B on tensor cores; returns DPP_ENOTSUP when alignment rules the path out
//   a_trans: A(m,k) = A[k*lda + m] else A[m*lda + k];  b_trans: B(n,k) = B[k*ldb + n] else B[n*ldb + k]
int dpp_gemm_tc(const float *A, const float *B, float *C, int M, int N, int K, int64_t lda, int64_t ldb, int a_trans,
                int b_trans, int precision, void *stream) {
    if (precision != 1 && precision != 2) return DPP_ENOTSUP;
    if ((lda & 3) || (ldb & 3) || (K & 3) || (M & 3) || (N & 3)) return DPP_ENOTSUP;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return DPP_ENOTSUP;
    GTArgs a{A, B, C, M, N, K, lda, ldb, 0, 0, 0};
    int rc;
    if (!a_trans && b_trans) rc = dispatch<false, true>(a, precision, S(stream));
    else if (!a_trans && !b_trans) rc = dispatch<false, false>(a, precision, S(stream));
    else if (a_trans && b_trans) rc = dispatch<true, true>(a, precision, S(stream));
    else return DPP_ENOTSUP;
    if (rc != 0) return dpp::fail(DPP_ECUDA, "%s: launch setup failed", __func__);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}
