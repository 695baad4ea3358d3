// Streaming 3xTF32 GEMM for the HiddenLayers (reference net/hiddenlayer.py:136-154 and its T.grad), written for the
// shape that matters: FC0 of the ResNet (16384 x 1024, net/resnet.py:141-145) at batch 128, where each of the three
// GEMMs of a step moves one 67 MB weight-sized matrix through HBM once and everything else is small.
//
//   D[m][n] = sum_k A(m,k) * B(n,k),   C[n*ldc + m] (+)= D[m][n]        (m = TMEM lane, 128 per tile)
//     forward  y  = x W      : m = output unit, n = sample, k = input unit    A = W  (m-contiguous), B = x    (k-contiguous)
//     backward dx = dpre W^T : m = input unit,  n = sample, k = output unit   A = W  (k-contiguous), B = dpre (k-contiguous)
//     backward dW = x^T dpre : m = output unit, n = input unit, k = sample    A = dpre (m-contiguous), B = x  (n-contiguous)
//   so the big matrix is always addressed with m (or n) along its rows exactly as it lies in memory and every global
//   store is a 128-byte row piece.
//
// One persistent CTA per SM, 14 warps:
//   warp 12 (one thread)  TMA: per 32-wide k-chunk two tensor loads (16 KB each) into a 4-deep landing ring.  A source
//                         that is k-contiguous arrives as [128 rows][32 k] with the hardware 128-byte swizzle, one
//                         that is m- / n-contiguous as [32 k][128 rows] unswizzled - both read conflict-free below.
//   warps 0-3             A transform: thread = row m; 32 k-values -> TF32 hi / lo -> tcgen05.st into its TMEM lane
//                         (the A operand of the MMAs lives in tensor memory: the weight stream never returns to
//                         shared memory after landing)
//   warps 4-7             B transform: thread = row n; 32 k-values -> hi / lo planes of a K-major SWIZZLE_128B tile
//   warp 13 (one thread)  12 x tcgen05.mma.kind::tf32 (.ts form, M = 128, N = bt, K = 8) per chunk, commits
//   warps 8-11            epilogue: tcgen05.ld, 128-byte row stores (or red.global.add when k is split over CTAs);
//                         accumulators are double-buffered so the epilogue of one unit hides under the next unit's MMAs
// Work units = (m-tile, n-tile, k-split); k is split only as far as needed to occupy the 148 SMs.
#include "tc_common.cuh"
#include <cuda.h>
#include <stdlib.h>

using namespace dpp;
using namespace dpp::tc;

namespace dpp {
struct FcEpilogue { const float *bias; const float *mask; float scale; int relu; };
}

namespace {

#ifdef DPP_PROFILE
// Debug timeline (tools/fc_stem_probe.py, PROBE_TIMELINE=1): CTA 0 appends (tag, clock64) pairs per role
__device__ long long *g_prof_fs = nullptr;
#define PROF_DECL(base_) int prof_n_ = (base_); long long *const prof_p_ = blockIdx.x == 0 ? g_prof_fs : nullptr
#define PROF(tag_)                                                                   \
    do {                                                                             \
        if (prof_p_ != nullptr && prof_n_ % 1000 < 990) {                            \
            prof_p_[prof_n_] = (tag_); prof_p_[prof_n_ + 1] = clock64(); prof_n_ += 2; \
        }                                                                            \
    } while (0)
#else
#define PROF_DECL(base_)
#define PROF(tag_)
#endif

constexpr int FS_THREADS = 448;
constexpr int NL = 5;                       // landing slots (32 KB each): as many as fit next to the operand tiles
constexpr int SLOT_BYTES = 32768;           // A part 16 KB | B part 16 KB
constexpr int BT_OFF = NL * SLOT_BYTES;     // B operand tiles: 2 stages x (hi 16 KB | lo 16 KB)
constexpr int BAR_OFF = BT_OFF + 2 * 32768;
constexpr int FS_SMEM = BAR_OFF + 256 + 1024;
constexpr int A_COL0 = 256;                 // TMEM: [0, 256) two accumulators, [256, 384) two A stages (hi 32 | lo 32 columns)

struct FSArgs {
    alignas(64) CUtensorMap tmA;
    alignas(64) CUtensorMap tmB;
    float *C;
    int64_t ldc;
    int M, N, K;
    int a_mn, b_mn;          // 1: the source is m- (n-) contiguous, landing tile [32 k][128 rows]; 0: k-contiguous, [128 rows][32 k]
    int bt;                  // MMA N: multiple of 16, <= 128
    int mtiles, ntiles, splits, chunks_per_split, total_chunks, units;
    int use_red;
    float *ws;               // k split over CTAs: unit u leaves its [128 n][128 m] partial tile at ws + u * 16384 (NULL: red.global.add)
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, uint64_t tmap, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}

struct Unit { int m0, n0, c_begin, c_end; };
__device__ __forceinline__ Unit decode_unit(const FSArgs &a, int u) {
    const int mt = u % a.mtiles, r = u / a.mtiles;
    const int nt = r % a.ntiles, sp = r / a.ntiles;
    Unit un;
    un.m0 = mt * 128; un.n0 = nt * 128;
    un.c_begin = sp * a.chunks_per_split;
    un.c_end = un.c_begin + a.chunks_per_split;
    if (un.c_end > a.total_chunks) un.c_end = a.total_chunks;
    return un;
}

// a row's 32 k-values out of a landing tile
__device__ __forceinline__ void read_row(const unsigned char *tile, int mn, int row, float (&v)[32]) {
    if (mn) {
        const float *p = reinterpret_cast<const float *>(tile) + row;
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = p[k * 128];
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 x = *reinterpret_cast<const float4 *>(tile + row * 128 + ((j ^ (row & 7)) << 4));
            v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
        }
    }
}

__global__ void __launch_bounds__(FS_THREADS, 1)
k_fc_stream(const __grid_constant__ FSArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // barriers: land_full[NL] 0.., land_empty[NL] NL.., a_full[2], b_full[2], st_empty[2], acc_full[2], acc_empty[2]
    auto bar = [&](int i) { return sbase + BAR_OFF + 8 * i; };
    constexpr int LF = 0, LE = NL, AF = 2 * NL, BF = AF + 2, SE = BF + 2, CF = SE + 2, CE = CF + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + BAR_OFF + 192);
    PROF_DECL(lane == 0 ? (warp == 0 ? 0 : warp == 4 ? 1000 : warp == 13 ? 2000 : warp == 8 ? 3000 : warp == 12 ? 4000 : 4990) : 4990);
    PROF(1);

    if (tid == 0) {
        for (int s = 0; s < NL; ++s) { mbar_init(bar(LF + s), 1); mbar_init(bar(LE + s), 8); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar(AF + s), 4); mbar_init(bar(BF + s), 4); mbar_init(bar(SE + s), 1);
            mbar_init(bar(CF + s), 1); mbar_init(bar(CE + s), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 13) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int gstride = gridDim.x;
    PROF(2);

    if (warp < 8) {
        // =========================== transform warps ===========================
        const bool isA = warp < 4;
        const int row = (warp & 3) * 32 + lane;
        const int mn = isA ? a.a_mn : a.b_mn;
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + A_COL0;
        const bool b_live = row < a.bt;
        uint32_t g = 0;
        for (int u = blockIdx.x; u < a.units; u += gstride) {
            const Unit un = decode_unit(a, u);
            for (int c = un.c_begin; c < un.c_end; ++c, ++g) {
                const uint32_t slot = g % NL, lph = (g / NL) & 1, stage = g & 1, sph = (g >> 1) & 1;
                if (lane == 0) mbar_wait_hint(bar(LF + slot), lph);
                __syncwarp();
                PROF(20);
                const unsigned char *tile = smem + slot * SLOT_BYTES + (isA ? 0 : 16384);
                float v[32];
                if (isA || b_live) read_row(tile, mn, row, v);
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar(LE + slot));        // the row is in registers: the landing slot can be refilled
                    mbar_wait_hint(bar(SE + stage), sph ^ 1);     // the MMAs that read this stage have retired
                }
                __syncwarp();
                PROF(21);
                if (isA) {
                    tc_fence_after();
                    const uint32_t ta = t_lane + stage * 64;
#pragma unroll
                    for (int hh = 0; hh < 4; ++hh) {
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int j = 0; j < 8; j += 2) split_tf32x2(v[hh * 8 + j], v[hh * 8 + j + 1], hi[j], hi[j + 1], lo[j], lo[j + 1]);
                        tmem_st8(ta + hh * 8, hi);
                        tmem_st8(ta + 32 + hh * 8, lo);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(AF + stage));
                    PROF(22);
                } else {
                    unsigned char *bt_tile = smem + BT_OFF + stage * 32768;
                    if (b_live) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int off = (row >> 3) * 1024 + (row & 7) * 128 + ((j ^ (row & 7)) << 4);
                            uint4 h, l;
                            split_tf32x2(v[4 * j], v[4 * j + 1], h.x, h.y, l.x, l.y);
                            split_tf32x2(v[4 * j + 2], v[4 * j + 3], h.z, h.w, l.z, l.w);
                            *reinterpret_cast<uint4 *>(bt_tile + off) = h;
                            *reinterpret_cast<uint4 *>(bt_tile + 16384 + off) = l;
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(BF + stage));
                    PROF(22);
                }
            }
        }
    } else if (warp == 12) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            const uint64_t tmA = reinterpret_cast<uint64_t>(&a.tmA), tmB = reinterpret_cast<uint64_t>(&a.tmB);
            uint32_t g = 0;
            for (int u = blockIdx.x; u < a.units; u += gstride) {
                const Unit un = decode_unit(a, u);
                for (int c = un.c_begin; c < un.c_end; ++c, ++g) {
                    const uint32_t slot = g % NL, lph = (g / NL) & 1;
                    mbar_wait_hint(bar(LE + slot), lph ^ 1);
                    PROF(10);
                    mbar_expect_tx(bar(LF + slot), SLOT_BYTES);
                    const uint32_t dst = sbase + slot * SLOT_BYTES;
                    // every CTA walks its k-range from a different starting chunk: CTAs running in lock step would
                    // otherwise all touch the same offset inside the (4 KB-strided) matrix rows at the same time
                    const int nch = un.c_end - un.c_begin;
                    int cr = c + (u * 5) % nch;
                    if (cr >= un.c_end) cr -= nch;
                    const int k0 = cr * 32;
                    if (a.a_mn) tma_load_2d(dst, tmA, un.m0, k0, bar(LF + slot)); else tma_load_2d(dst, tmA, k0, un.m0, bar(LF + slot));
                    if (a.b_mn) tma_load_2d(dst + 16384, tmB, un.n0, k0, bar(LF + slot)); else tma_load_2d(dst + 16384, tmB, k0, un.n0, bar(LF + slot));
                }
            }
        }
        __syncwarp();
    } else if (warp == 13) {
        // =========================== MMA issuer ===========================
        if (elect_one()) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.bt >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            uint32_t g = 0, j = 0;
            for (int u = blockIdx.x; u < a.units; u += gstride, ++j) {
                const Unit un = decode_unit(a, u);
                const uint32_t acc = j & 1, aph = (j >> 1) & 1;
                mbar_wait_hint(bar(CE + acc), aph ^ 1);
                tc_fence_after();
                PROF(40);
                const uint32_t d_tmem = tmem_base + acc * 128;
                for (int c = un.c_begin; c < un.c_end; ++c, ++g) {
                    const uint32_t stage = g & 1, sph = (g >> 1) & 1;
                    mbar_wait_hint(bar(AF + stage), sph);
                    PROF(41);
                    mbar_wait_hint(bar(BF + stage), sph);
                    tc_fence_after();
                    PROF(42);
                    const uint32_t ta = tmem_base + A_COL0 + stage * 64;
                    const uint64_t b0 = make_desc(sbase + BT_OFF + stage * 32768);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t bh = b0 + 2 * ks, bl = bh + (16384 >> 4);
                        const uint32_t first = (c == un.c_begin && ks == 0) ? 0u : 1u;
                        mma_tf32_ts_1t(d_tmem, ta + ks * 8, bl, idesc, first);
                        mma_tf32_ts_1t(d_tmem, ta + 32 + ks * 8, bh, idesc, 1u);
                        mma_tf32_ts_1t(d_tmem, ta + ks * 8, bh, idesc, 1u);
                    }
                    mma_commit_1t(bar(SE + stage));
                    if (c == un.c_end - 1) mma_commit_1t(bar(CF + acc));
                    PROF(43);
                }
            }
        }
        __syncwarp();
    } else {
        // =========================== epilogue ===========================
        const int q = warp - 8;
        uint32_t j = 0;
        for (int u = blockIdx.x; u < a.units; u += gstride, ++j) {
            const Unit un = decode_unit(a, u);
            const uint32_t acc = j & 1, aph = (j >> 1) & 1;
            if (lane == 0) mbar_wait_hint(bar(CF + acc), aph);
            __syncwarp();
            tc_fence_after();
            PROF(50);
            const int m = un.m0 + q * 32 + lane;
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 128;
            const int rot = ((u * 3) % (a.bt >> 4)) << 4;        // ... and its output columns from a different one
            if (a.ws != nullptr) {
                float *dst = a.ws + (size_t)u * 16384 + q * 32 + lane;
                for (int cb0 = 0; cb0 < a.bt; cb0 += 16) {
                    const int cb = cb0 + rot < a.bt ? cb0 + rot : cb0 + rot - a.bt;
                    float v[16];
                    tmem_ld16(t0 + cb, v);
                    PROF(52);
#pragma unroll
                    for (int t = 0; t < 16; ++t) dst[(cb + t) * 128] = v[t];
                    PROF(53);
                }
            } else {
                // interior tiles (the common case) take a branch-free path: 16 independent row stores per TMEM load
                const bool full = un.n0 + a.bt <= a.N;
                const bool red = a.use_red != 0;
                const size_t ldc = (size_t)a.ldc;
                for (int cb0 = 0; cb0 < a.bt; cb0 += 16) {
                    const int cb = cb0 + rot < a.bt ? cb0 + rot : cb0 + rot - a.bt;
                    float v[16];
                    tmem_ld16(t0 + cb, v);
                    PROF(52);
                    if (m < a.M) {
                        float *dst = a.C + (size_t)(un.n0 + cb) * ldc + m;
                        if (full && !red) {
#pragma unroll
                            for (int t = 0; t < 16; ++t) dst[t * ldc] = v[t];
                        } else if (full) {
#pragma unroll
                            for (int t = 0; t < 16; ++t) atomicAdd(dst + t * ldc, v[t]);
                        } else {
                            const int nleft = a.N - (un.n0 + cb);
#pragma unroll
                            for (int t = 0; t < 16; ++t)
                                if (t < nleft) {
                                    if (red) atomicAdd(dst + t * ldc, v[t]); else dst[t * ldc] = v[t];
                                }
                        }
                    }
                    PROF(53);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(CE + acc));
            PROF(51);
        }
    }
    PROF(3);
    tc_fence_before();
    __syncthreads();
    if (warp == 13) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// Sum of the k-split partial tiles in split order (deterministic), with the HiddenLayer epilogue when a bias is given:
// y = act(sum + bias) * mask * scale  (net/hiddenlayer.py:136-154, net/dropoutlayer.py:104)
__global__ void __launch_bounds__(256)
k_fs_reduce(const float *__restrict__ ws, float *__restrict__ C, int64_t ldc, int M, int N, int mtiles, int ntiles, int splits,
            int add, const float *__restrict__ bias, const float *__restrict__ mask, float scale, int relu) {
    const int64_t total = (int64_t)M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i % M), n = (int)(i / M);
        const float *p = ws + ((size_t)(n >> 7) * mtiles + (m >> 7)) * 16384 + (n & 127) * 128 + (m & 127);
        const size_t sstride = (size_t)ntiles * mtiles * 16384;
        float s = 0.f;
        for (int sp = 0; sp < splits; ++sp) s += p[sp * sstride];
        float *c = C + (size_t)n * ldc + m;
        if (bias != nullptr) {
            s += bias[m];
            if (relu) s = fmaxf(s, 0.f);
            if (mask != nullptr) s *= mask[(size_t)n * ldc + m];
            s *= scale;
        }
        *c = add ? *c + s : s;
    }
}

float *g_fs_ws = nullptr;
constexpr size_t FS_WS_FLOATS = (size_t)148 * 16384;
// the partial-tile workspace (9.7 MB) is library-owned; it cannot be allocated inside a stream capture
float *fs_workspace(cudaStream_t st) {
    if (g_fs_ws == nullptr) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (st != nullptr && (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone)) return nullptr;
        if (cudaMalloc(&g_fs_ws, FS_WS_FLOATS * sizeof(float)) != cudaSuccess) { g_fs_ws = nullptr; cudaGetLastError(); }
    }
    return g_fs_ws;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn fs_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D map of a row-major matrix [rows][ld]: `mn` = 1 -> the operand's m (n) index runs along the memory rows' elements
// (box 128 x 32 k-rows, unswizzled); 0 -> k runs along them (box 32 k x 128 rows, SWIZZLE_128B)
int make_map(CUtensorMap *tm, const float *p, int mn, int extent_rows_mn, int extent_k, int64_t ld) {
    EncodeTiledFn enc = fs_encode_tiled();
    if (enc == nullptr) return -1;
    cuuint64_t gdim[2], gstr[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2], estr[2] = {1, 1};
    if (mn) { gdim[0] = (cuuint64_t)extent_rows_mn; gdim[1] = (cuuint64_t)extent_k; box[0] = 128; box[1] = 32; }
    else { gdim[0] = (cuuint64_t)extent_k; gdim[1] = (cuuint64_t)extent_rows_mn; box[0] = 32; box[1] = 128; }
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(p), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, mn ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

}  // namespace

namespace dpp {

int fc_stream_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("DPP_FC_STREAM"); v = e ? atoi(e) : 1; }
    return v;
}

// C[n*ldc + m] (+)= sum_k A(m,k) B(n,k) in 3xTF32.  a_mn: A(m,k) = A[k*lda + m] else A[m*lda + k]; b_mn likewise for B(n,k).
// add = 1: the result is added to C; add = 0: C is overwritten.  When k has to be split over CTAs the partial tiles go
// through the library workspace and a second small kernel sums them in split order; `epi` (may be NULL; requires add = 0)
// = the HiddenLayer epilogue, applied by that kernel - *epi_done tells the caller whether it was.  Returns DPP_ENOTSUP
// when the layout rules TMA out.
int fc_stream_gemm(const float *A, int a_mn, int64_t lda, const float *B, int b_mn, int64_t ldb, float *C, int64_t ldc, int M,
                   int N, int K, int add, const FcEpilogue *epi, int *epi_done, cudaStream_t st) {
    if (epi_done) *epi_done = 0;
    if (!fc_stream_enabled()) return DPP_ENOTSUP;
    if ((lda & 3) || (ldb & 3) || M < 1 || N < 1 || K < 1) return DPP_ENOTSUP;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return DPP_ENOTSUP;
    FSArgs a;
    memset(&a, 0, sizeof(a));
    if (make_map(&a.tmA, A, a_mn, M, K, lda) != 0) return DPP_ENOTSUP;
    if (make_map(&a.tmB, B, b_mn, N, K, ldb) != 0) return DPP_ENOTSUP;
    a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K; a.a_mn = a_mn; a.b_mn = b_mn;
    a.bt = N >= 128 ? 128 : ((N + 15) / 16) * 16;
    a.mtiles = (M + 127) / 128; a.ntiles = (N + 127) / 128;
    a.total_chunks = (K + 31) / 32;
    const int base = a.mtiles * a.ntiles;
    int splits = base >= 148 ? 1 : 148 / base;
    if (splits > a.total_chunks / 2) splits = a.total_chunks / 2 > 0 ? a.total_chunks / 2 : 1;
    a.chunks_per_split = (a.total_chunks + splits - 1) / splits;
    a.splits = (a.total_chunks + a.chunks_per_split - 1) / a.chunks_per_split;
    a.units = base * a.splits;
    a.ws = a.splits > 1 ? fs_workspace(st) : nullptr;          // units <= 148 whenever k is split
    a.use_red = (a.ws == nullptr && (add || a.splits > 1)) ? 1 : 0;
    if (a.use_red && !add) {
        if (ldc == M) { if (cudaMemsetAsync(C, 0, sizeof(float) * (size_t)N * M, st) != cudaSuccess) return DPP_ECUDA; }
        else if (cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * M, N, st) != cudaSuccess) return DPP_ECUDA;
    }
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(k_fc_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM) != cudaSuccess) return DPP_ECUDA;
        attr_done = true;
    }
    const int grid = a.units < 148 ? a.units : 148;
    k_fc_stream<<<grid, FS_THREADS, FS_SMEM, st>>>(a);
    if (cudaGetLastError() != cudaSuccess) return DPP_ECUDA;
    if (a.ws != nullptr) {
        const int64_t total = (int64_t)M * N;
        const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
        const bool e = epi != nullptr && !add;
        k_fs_reduce<<<blocks, 256, 0, st>>>(a.ws, C, ldc, M, N, a.mtiles, a.ntiles, a.splits, add, e ? epi->bias : nullptr,
                                            e ? epi->mask : nullptr, e ? epi->scale : 1.f, e ? epi->relu : 0);
        if (cudaGetLastError() != cudaSuccess) return DPP_ECUDA;
        if (e && epi_done) *epi_done = 1;
    }
    return DPP_OK;
}

}  // namespace dpp

// Allocates the library-owned workspace of the k-split HiddenLayer GEMMs (9.7 MB, idempotent).  Call once outside CUDA-graph
// capture; without it a captured dpp_fc_fwd / dpp_fc_bwd falls back to red.global.add reductions.
extern "C" int dpp_fc_workspace_init(void) {
    if (fs_workspace(nullptr) == nullptr) return dpp::fail(DPP_ECUDA, "%s: workspace allocation failed", __func__);
    return DPP_OK;
}

#ifdef DPP_PROFILE
extern "C" int dpp_debug_set_prof_fs(void *buf) {
    long long *p = reinterpret_cast<long long *>(buf);
    DPP_CUDA(cudaMemcpyToSymbol(g_prof_fs, &p, sizeof(p)));
    return DPP_OK;
}
#endif
