// tcgen05 implicit-GEMM convolution for sm_100a: forward and backward-data of every ConvLayer of the
// ResNet (reference net/convlayer.py:230-235; 1x1, 1x1/s2, 3x3 'half'), with the BatchNorm+ReLU
// prologue on the gathered operand and the bias / residual / BN-statistics (forward) or ReLU-mask /
// BN-backward-statistics (dgrad) epilogue fused in - the same fusion contract as conv_simt.cu.
//
// Precision modes: 2 = TF32 (one MMA pass), 1 = 3xTF32 (x = hi + lo with hi = rna_tf32(x),
// lo = rna_tf32(x - hi); D += Ahi*Blo + Alo*Bhi + Ahi*Bhi) which recovers fp32-level products and is
// the mode that meets the 1e-4 parity bar.  Accumulation is fp32 in TMEM.
//
// GEMM view per CTA tile: D[128 pixels][BN channels] += A[128][32] * B[BN][32]^T per k-chunk,
//   A: gathered from NHWC global memory by 128 producer threads (one pixel row each: up to 128
//      contiguous bytes per tap), BN+ReLU applied in registers, split hi/lo, written to shared
//      memory in the UMMA K-major SWIZZLE_128B layout (generic-proxy stores + fence.proxy.async);
//   B: the layer's weights, pre-packed once per step by dpp_conv_pack_all into the exact
//      shared-memory image (hi/lo, swizzled) of every (n-tile, k-chunk): ONE cp.async.bulk (TMA
//      bulk copy, UBLKCP) per stage, completing on the stage's mbarrier;
//   MMA: a single thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8) with the
//      accumulator in TMEM (double-buffered: 2*BN columns) and commits to mbarriers;
//   epilogue: 4 warps read TMEM with tcgen05.ld.32x32b, apply the fused epilogue and store NHWC rows.
// Warp roles: warps 0-3 producers, warps 4-7 epilogue (TMEM lane quarter = warp%4), warp 8 MMA issuer +
// TMEM allocator.  Persistent CTAs (<= 1 per SM) walk the tile list, so per-channel fp64 statistics
// leave the CTA once.
#include "tc_common.cuh"
#include <cuda.h>
#include <stdlib.h>

using namespace dpp;
using namespace dpp::tc;

namespace {

#ifdef DPP_PROFILE
// Debug timeline (tools/conv_probe.py): CTA 0 appends (tag, clock64) pairs per role into a global buffer.
__device__ long long *g_prof = nullptr;
__device__ int g_dbg = 0;     // experiment knobs: 1 no global loads, 2 no transform/STTM, 4 no MMA, 8 no epilogue body, 16 no setup math
#define DBG(bit_) (g_dbgv & (bit_))
#define DBG_DECL const int g_dbgv = g_dbg
#define PROF_DECL(base_) int prof_n_ = (base_); long long *const prof_p_ = blockIdx.x == 0 ? g_prof : nullptr
#define PROF(tag_)                                                                   \
    do {                                                                             \
        if (prof_p_ != nullptr && prof_n_ % 1000 < 990) {                            \
            prof_p_[prof_n_] = (tag_); prof_p_[prof_n_ + 1] = clock64(); prof_n_ += 2; \
        }                                                                            \
    } while (0)
#else
#define PROF_DECL(base_)
#define PROF(tag_)
#define DBG(bit_) false
#define DBG_DECL
#endif

constexpr int TM = 128;          // pixels per tile (TMEM lanes)
constexpr int KC = 32;           // floats of K per stage: 128-byte rows
constexpr int NTHREADS_CONV = 576;  // k_conv_tc: 8 producer + 2 x 4 epilogue warps, MMA issuer, weight-image (TMA) loader

struct TCArgs {
    // TMA descriptor of the gathered tensor: rank 4 {C, W, H, N}, box = {min(Cin, 32) channels, the tile's pixel
    // rectangle}, traversal strides = the convolution stride, zero fill outside the image ('half' padding for free)
    alignas(64) CUtensorMap tmap;
    int box_h, box_n;     // the 128 GEMM rows of a tile = box_n images x box_h rows x Wg pixels
    const float *in;      // gathered tensor [N, Hin, Win, Cin]
    const float *wimg;    // packed weight image for this mode
    float *out;           // [N, Hout, Wout, Cn]
    int N, Hin, Win, Cin;
    int Hg, Wg;
    int Hout, Wout, Cn;
    int k, pad, in_stride, out_stride;
    int wmode;            // 0 forward, 1 dgrad
    int kchunks;          // ceil(k*k*Cin / 32)
    int wsh, hsh;         // log2(Wg), log2(Hg) when both are powers of two, else -1 (generic division)
    int knobs;            // tuning bits (DPP_TC_KNOBS): 4 = skip the statistics atomics (timing ablation)
    dpp_bn_ref in_bn; int has_in_bn;
    const float *bias; const float *residual; double *out_stats;
    int accumulate; dpp_bn_ref mask_bn; int has_mask; const float *x_pre; double *dz_stats;
    // fused BatchNorm backward behind a grid-wide barrier (dgrad of the LAST consumer of a normalised tensor):
    // after every CTA has written its dz tiles and added its statistics, dx = bn_bwd(dz, x_pre) [+ skip]
    int tail; const float *tail_skip; float *tail_out; float *dgamma; float *dbeta; float pscale; unsigned int *gbar;
    // per k-chunk gather table (host-built): the chunk's 8 16-byte pieces come from tap A (pieces 0-3) and
    // tap B (pieces 4-7; same tap as A when Cin >= 32): {dr, ds (tap offset incl. -pad), element offset, channel}
    int tab[18][8];
};

// Epilogue geometry: the accumulator tile leaves TMEM row-per-thread, is transposed through a per-warp
// shared staging tile in batches of CB columns, and is then handled column-per-lane-group so that global
// loads/stores are 16-byte pieces of contiguous NHWC rows (coalesced) and column statistics are plain
// per-thread sums.
template <int BN>
struct EpiGeo {
    static constexpr int CB = 16;                     // columns per batch (64-byte row segments)
    static constexpr int LPR = CB / 4;                // lanes per row (16 B each)
    static constexpr int RPI = 32 / LPR;              // rows per warp instruction
    static constexpr int NIT = 32 / RPI;              // instructions per batch
    static constexpr int NB = BN / CB;                // batches per tile
    static constexpr int NSTEP = NB * NIT;
    static constexpr int PD = NSTEP < 8 ? NSTEP : 8;  // side-operand prefetch distance (steps)
    static constexpr int ROW_BYTES = CB == 32 ? 144 : 64;
    static constexpr int WARP_BYTES = 32 * ROW_BYTES;
    __device__ static __forceinline__ uint32_t addr(int row, int c4) {
        return CB == 32 ? (uint32_t)(row * 144 + c4 * 16) : (uint32_t)(row * 64 + ((c4 ^ ((row >> 1) & 3)) << 4));
    }
};

// smem carve-up (after 1024-byte alignment):
//   B ring [RB][PASSES][BN rows][128 B] (weight images, TMA bulk copies), barriers, epilogue staging tiles,
//   BN coefficients, chunk table, per-warp fp64 statistics.  The A operand lives in TENSOR MEMORY.
// TMEM columns: [0, 2*BN) two accumulators, [256 + s*32*PASSES, ...) A stage s (hi: K = 32 columns, lo: next 32).
template <int BN, int PASSES>
struct SmemLayout {
    static constexpr int B_BYTES = PASSES * BN * 128;
    static constexpr int NS = 4;                                   // A stages (TMEM)
    static constexpr int RB = (B_BYTES >= 32768) ? 3 : (B_BYTES >= 16384 ? 4 : 6);   // weight-image ring slots
    static constexpr int NSLOT = (B_BYTES >= 32768) ? 4 : 6;       // landing slots of the gathered operand (16 KB k-chunks, TMA)
    static constexpr int A_COL0 = 256, A_COLS = 32 * PASSES;
    static constexpr int B_OFF = 0;
    static constexpr int RAW_OFF = B_OFF + RB * B_BYTES;           // multiple of 1024: the swizzled TMA boxes land here
    static constexpr int BAR_OFF = RAW_OFF + NSLOT * TM * 128;
    static constexpr int STG_OFF = BAR_OFF + 512;
    static constexpr int COEF_OFF = STG_OFF + 8 * EpiGeo<BN>::WARP_BYTES;
    static constexpr int COEF_BYTES = 2 * 256 * 4 + 5 * 128 * 4 + 18 * 32;   // in-BN scale/shift, n-tile coefficients, chunk table
    static constexpr int STAT_OFF = COEF_OFF + COEF_BYTES;
    static constexpr int TOTAL = STAT_OFF + 8 * 2 * BN * 8 + 1024;
};

constexpr int NPROD_WARPS = 8;
constexpr int W_EPI = 8, W_MMA = 16, W_LOAD = 17;      // warps 8-11 / 12-15: epilogue warpgroups 0 / 1

// decode a GEMM row (pixel of the gather grid) -> (n, ho, wo)
__device__ __forceinline__ void decode_pix(const TCArgs &a, int m, int &n, int &ho, int &wo) {
    if (a.wsh >= 0) {
        wo = m & (a.Wg - 1); ho = (m >> a.wsh) & (a.Hg - 1); n = m >> (a.wsh + a.hsh);
    } else {
        wo = m % a.Wg; ho = (m / a.Wg) % a.Hg; n = m / (a.Wg * a.Hg);
    }
}

template <int BN, int PASSES>
__global__ void __launch_bounds__(NTHREADS_CONV, 1)
k_conv_tc(const __grid_constant__ TCArgs a) {
    using L = SmemLayout<BN, PASSES>;
    using G = EpiGeo<BN>;
    constexpr int NS = L::NS, RB = L::RB;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still in the shared window
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    PROF_DECL(lane == 0 ? (warp == 0 ? 0 : warp == W_MMA ? 1000 : warp == W_EPI ? 2000 : warp == W_LOAD ? 3000 : 4000) : 4000);
    PROF(1);
    DBG_DECL;
    // bar index: full[s] = s, empty[s] = NS + s, tfull[a] = 2*NS + a, tempty[a] = 2*NS + 2 + a,
    //            bfull[b] = 2*NS + 4 + b, bempty[b] = 2*NS + 4 + RB + b,
    //            rfull[r] = RBAR + r, rempty[r] = RBAR + NSLOT + r   (landing ring of the gathered operand)
    constexpr int NSLOT = L::NSLOT, RBAR = 2 * NS + 4 + 2 * RB;
    static_assert(8 * (RBAR + 2 * NSLOT) <= 480, "barrier region");
    auto bar = [&](int i) { return sbase + L::BAR_OFF + 8 * i; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::BAR_OFF + 480);
    float *s_scale = reinterpret_cast<float *>(smem + L::COEF_OFF);   // [256] input-BN scale
    float *s_shift = s_scale + 256;              // [256]
    float *s_msc = s_shift + 256;                // [BN] mask-BN scale   (this CTA's n-tile)
    float *s_msh = s_msc + 128;                  // [BN] mask-BN shift
    float *s_mmean = s_msh + 128;                // [BN]
    float *s_mistd = s_mmean + 128;              // [BN]
    float *s_bias = s_mistd + 128;               // [BN] forward bias of this CTA's n-tile
    int4 *s_tab = reinterpret_cast<int4 *>(s_bias + 128);   // [18][2] chunk gather table
    double *s_stat = reinterpret_cast<double *>(smem + L::STAT_OFF);   // [8 warps][2 kinds][BN]

    const int M = a.N * a.Hg * a.Wg;
    const int mtiles = (M + TM - 1) / TM;
    const int ntiles = a.Cn / BN;
    const int tiles = mtiles * ntiles;
    // gridDim.x is a multiple of ntiles, so every tile of this CTA has the same n-tile
    const int cta_n0 = (blockIdx.x % ntiles) * BN;
    constexpr uint32_t TCOLS = 512;              // 2 accumulators + the A stages; one CTA per SM
    const bool resident = a.kchunks <= RB;       // the whole weight image of this n-tile stays in the ring

    // ---- private set-up: nothing here touches global memory, so under programmatic dependent launch it
    //      overlaps the tail of the previous kernel of the chain
    pdl_trigger();
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(bar(s), NPROD_WARPS / 2); mbar_init(bar(NS + s), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar(2 * NS + i), 1); mbar_init(bar(2 * NS + 2 + i), 4); }
        for (int b = 0; b < RB; ++b) { mbar_init(bar(2 * NS + 4 + b), 1); mbar_init(bar(2 * NS + 4 + RB + b), 1); }
        for (int r = 0; r < NSLOT; ++r) { mbar_init(bar(RBAR + r), 1); mbar_init(bar(RBAR + NSLOT + r), NPROD_WARPS / 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int c = tid; c < a.kchunks * 2; c += NTHREADS_CONV)
        s_tab[c] = make_int4(a.tab[c >> 1][(c & 1) * 4], a.tab[c >> 1][(c & 1) * 4 + 1], a.tab[c >> 1][(c & 1) * 4 + 2], a.tab[c >> 1][(c & 1) * 4 + 3]);
    for (int c = tid; c < 8 * 2 * BN; c += NTHREADS_CONV) s_stat[c] = 0.0;
    __syncthreads();                             // barrier inits visible to every role (the loader starts right after the wait)
    // ---- everything below reads what the previous kernels wrote (statistics, activations, weight images)
    pdl_wait();
    // The loader warp goes straight to its role: its first TMA round trip (~1 us) overlaps the coefficient phase of the
    // other warps (statistics from L2 + fp64 arithmetic, ~1 us) instead of following it at the head of every launch of
    // the chain.  The other 17 warps meet at a named barrier once the coefficients are in shared memory.
    if (warp != W_LOAD) {
    // Three independent coefficient jobs on disjoint thread ranges, every job with all its global loads issued
    // before the fp64 arithmetic: one exposed memory latency instead of a chain of dependent ones.
    if (tid < 256) {                                   // input-BN scale / shift (Cin <= 256)
        const int c = tid;
        if (a.has_in_bn && c < a.Cin) {
            const bool batch = a.in_bn.sums != nullptr;
            const double s1 = batch ? a.in_bn.sums[c] : 0.0, s2 = batch ? a.in_bn.sums[a.Cin + c] : 0.0;
            const float rm = batch ? 0.f : a.in_bn.mean[c], ri = batch ? 0.f : a.in_bn.inv_std[c];
            const float g = a.in_bn.gamma[c], be = a.in_bn.beta[c];
            float mean = rm, istd = ri;
            if (batch) {
                const double m = s1 / a.in_bn.count;
                double var = s2 / a.in_bn.count - m * m;
                if (var < 0.0) var = 0.0;
                mean = (float)m;
                istd = (float)(1.0 / sqrt(var + (double)a.in_bn.eps));
            }
            const float sc = g * istd;
            s_scale[c] = sc; s_shift[c] = be - mean * sc;
        }
    } else if (tid < 256 + 128) {                      // mask-BN coefficients of this CTA's n-tile (dgrad)
        const int c = tid - 256;
        if (a.has_mask && c < BN) {
            const int cg = cta_n0 + c;
            const bool batch = a.mask_bn.sums != nullptr;
            const double s1 = batch ? a.mask_bn.sums[cg] : 0.0, s2 = batch ? a.mask_bn.sums[a.Cn + cg] : 0.0;
            const float rm = batch ? 0.f : a.mask_bn.mean[cg], ri = batch ? 0.f : a.mask_bn.inv_std[cg];
            const float g = a.mask_bn.gamma[cg], be = a.mask_bn.beta[cg];
            float mean = rm, istd = ri;
            if (batch) {
                const double m = s1 / a.mask_bn.count;
                double var = s2 / a.mask_bn.count - m * m;
                if (var < 0.0) var = 0.0;
                mean = (float)m;
                istd = (float)(1.0 / sqrt(var + (double)a.mask_bn.eps));
            }
            const float sc = g * istd;
            s_msc[c] = sc; s_msh[c] = be - mean * sc;
            s_mmean[c] = mean; s_mistd[c] = istd;
        }
    } else if (tid < 256 + 256) {                      // forward bias of this CTA's n-tile
        const int c = tid - 384;
        if (c < BN) s_bias[c] = (a.wmode == 0 && a.bias) ? a.bias[cta_n0 + c] : 0.f;
    }
    tc_fence_before();
    asm volatile("bar.sync 2, %0;" ::"n"(NTHREADS_CONV - 32) : "memory");
    tc_fence_after();
    }
    const uint32_t tmem_base = *tmem_slot;       // (the loader warp may read this before the allocation: it never uses it)
    const int my_tiles = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA
    PROF(2);

    if (warp < NPROD_WARPS) {
        // =========================== transform warps ===========================
        // Two independent groups of 4 warps; group g owns the k-chunks with (chunk index % 2) == g, so the
        // synchronisation latencies of the two groups overlap.  warp % 4 = TMEM lane quarter; a thread owns one GEMM
        // row (pixel).  The loader warp stages every chunk - the tile's 128 pixels x 32 K-values of one or two filter
        // taps - with TMA tensor loads (cp.async.bulk.tensor, zero fill outside the image) into a landing ring.  A
        // thread reads its row (8 x 16 bytes, hardware-swizzled: conflict-free), applies BN + ReLU, splits into TF32
        // hi / lo and writes 32 + 32 columns of its TMEM lane (tcgen05.st): the MMA reads A from tensor memory.
        // No address arithmetic and no load instructions per element are left in these warps (round 1 gathered
        // with 1024 cp.async per chunk, a third of the chunk period).
        const int q = warp & 3, grp = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + L::A_COL0;
        const int kchunks = a.kchunks, Hin = a.Hin, Win = a.Win, gstride = gridDim.x;
        const int T = my_tiles * kchunks;
        const int Tg = (T - grp + 1) / 2;             // chunks of this group: grp, grp + 2, ...
        const bool pro = a.has_in_bn != 0, relu = a.in_bn.relu != 0;
        // where the row's 8 pieces lie in a landing slot.  Cin >= 32: one box of 32 channels, 128-byte rows,
        // SWIZZLE_128B (piece ^ (row & 7)).  Cin == 16: two boxes of 16 channels (tap A, tap B) of 8 KB each, 64-byte
        // rows, SWIZZLE_64B (piece ^ ((row >> 1) & 3)).
        const bool c16 = a.Cin == 16;
        uint32_t poff[8];
#pragma unroll
        for (int pj = 0; pj < 8; ++pj)
            poff[pj] = c16 ? (uint32_t)((pj >> 2) * 8192 + row * 64 + (((pj & 3) ^ ((row >> 1) & 3)) << 4))
                           : (uint32_t)(row * 128 + ((pj ^ (row & 7)) << 4));
        int i_tile = blockIdx.x, i_kc = grp;
        int r_h0, r_w0; bool r_ok;    // this thread's own row: top-left input pixel, row inside M
        auto set_tile = [&]() {
            const int m = (i_tile / ntiles) * TM + row;
            int n, ho, wo;
            decode_pix(a, m < M ? m : 0, n, ho, wo);
            r_h0 = ho * a.in_stride; r_w0 = wo * a.in_stride;
            r_ok = m < M;
        };
        while (i_kc >= kchunks) { i_kc -= kchunks; i_tile += gstride; }
        set_tile();
        uint32_t stage = grp, phase = 0;              // A stage in TMEM (NS stages, this group uses grp, grp + 2)
        uint32_t slot = grp, sphase = 0;              // landing slot (NSLOT slots, this group uses grp, grp + 2, ...)
#pragma unroll 1
        for (int i = 0; i < Tg; ++i) {
            const int kc = i_kc;
            const int4 e0 = s_tab[kc * 2], e1 = s_tab[kc * 2 + 1];
            // taps outside the image (and K-values beyond k*k*Cin) are zero AFTER BN + ReLU
            const bool v0 = r_ok && (unsigned)(r_h0 + e0.x) < (unsigned)Hin && (unsigned)(r_w0 + e0.y) < (unsigned)Win;
            const bool v1 = r_ok && (unsigned)(r_h0 + e1.x) < (unsigned)Hin && (unsigned)(r_w0 + e1.y) < (unsigned)Win;
            i_kc += 2;
            if (i_kc >= kchunks) {
                do { i_kc -= kchunks; i_tile += gstride; } while (i_kc >= kchunks);
                set_tile();
            }
            if (lane == 0) mbar_wait(bar(RBAR + slot), sphase);       // the chunk has landed
            __syncwarp();
            PROF(11);
            const unsigned char *rp = smem + L::RAW_OFF + slot * (TM * 128);
            float4 xr[8];
#pragma unroll
            for (int pj = 0; pj < 8; ++pj) xr[pj] = *reinterpret_cast<const float4 *>(rp + poff[pj]);
#pragma unroll
            for (int pj = 0; pj < 8; ++pj) {
                float4 x = xr[pj];
                if (pro) {
                    const int chan = (pj < 4 ? e0.w : e1.w) + (pj & 3) * 4;
                    const float4 sc = *reinterpret_cast<const float4 *>(s_scale + chan);
                    const float4 sf = *reinterpret_cast<const float4 *>(s_shift + chan);
                    const float2 p0 = __ffma2_rn(make_float2(x.x, x.y), make_float2(sc.x, sc.y), make_float2(sf.x, sf.y));
                    const float2 p1 = __ffma2_rn(make_float2(x.z, x.w), make_float2(sc.z, sc.w), make_float2(sf.z, sf.w));
                    x = make_float4(p0.x, p0.y, p1.x, p1.y);         // (packed FMAs: same roundings, half the issue slots)
                    if (relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                }
                xr[pj] = x;
            }
            if (!__all_sync(0xffffffffu, v0 && v1)) {      // only warps that touch the image border (or the tail of M) pay for the zeroing
#pragma unroll
                for (int pj = 0; pj < 8; ++pj)
                    if (!(pj < 4 ? v0 : v1)) xr[pj] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (lane == 0) mbar_wait(bar(NS + stage), phase ^ 1);     // the MMAs that read this A stage have retired
            __syncwarp();
            tc_fence_after();
            PROF(12);
            const uint32_t ta = t_lane + stage * L::A_COLS;
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) {           // 8 columns (two 16-byte pieces) at a time
                if (DBG(2)) break;
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float4 x = xr[hh * 2 + j];   // piece 0-3: tap A, 4-7: tap B
                    split_tf32x2(x.x, x.y, hi[4 * j], hi[4 * j + 1], lo[4 * j], lo[4 * j + 1]);
                    split_tf32x2(x.z, x.w, hi[4 * j + 2], hi[4 * j + 3], lo[4 * j + 2], lo[4 * j + 3]);
                }
                tmem_st8(ta + hh * 8, hi);
                if (PASSES > 1) tmem_st8(ta + 32 + hh * 8, lo);
            }
            tmem_st_wait();
            PROF(14);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar(stage));                      // A stage written
                mbar_arrive(bar(RBAR + NSLOT + slot));        // landing slot consumed (every lane's row went through registers)
            }
            PROF(13);
            stage += 2;
            if (stage >= NS) { stage -= NS; phase ^= 1; }
            slot += 2;
            if (slot >= NSLOT) { slot -= NSLOT; sphase ^= 1; }
        }
    } else if (warp == W_MMA) {
        // =========================== MMA issuer ===========================
        // ONE elected thread runs the whole role (waits, tcgen05.mma, commits): inside a single-thread region ptxas keeps
        // descriptors and tensor-memory addresses in uniform registers and issues the 12 MMAs of a k-chunk back to back
        // (the per-instruction elect.sync of round 1 cost ~55 cycles per MMA: ELECT / VOTEU / R2UR chains).
        if (elect_one()) {
            constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            constexpr uint32_t LO_OFF = (BN * 128) >> 4;        // descriptor units (16 B) from the hi to the lo weight plane
            uint32_t acc = 0, aphase = 0;
            uint32_t stage = 0, phase = 0, bslot = 0, bphase = 0;
            const int kchunks = a.kchunks;
            for (int t = 0; t < my_tiles; ++t) {
                mbar_wait(bar(2 * NS + 2 + acc), aphase ^ 1);
                tc_fence_after();
                PROF(20);
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kc = 0; kc < kchunks; ++kc) {
                    if (resident) { bslot = kc; bphase = 0; }
                    if (!resident || t == 0) mbar_wait(bar(2 * NS + 4 + bslot), bphase);   // resident images are complete after the first tile
                    PROF(21);
                    mbar_wait(bar(stage), phase);
                    tc_fence_after();
                    PROF(22);
                    const uint32_t ta = tmem_base + L::A_COL0 + stage * L::A_COLS;
                    const uint64_t b0 = make_desc(sbase + L::B_OFF + bslot * L::B_BYTES);
#pragma unroll
                    for (int ks = 0; ks < KC / 8; ++ks) {
                        if (DBG(4)) break;
                        const uint64_t bh = b0 + 2 * ks;       // + ks * 32 bytes along K inside the swizzled row
                        const uint32_t first = (kc == 0 && ks == 0) ? 0u : 1u;
                        if (PASSES > 1) {
                            const uint64_t bl = bh + LO_OFF;
                            mma_tf32_ts_1t(d_tmem, ta + ks * 8, bl, IDESC, first);
                            mma_tf32_ts_1t(d_tmem, ta + 32 + ks * 8, bh, IDESC, 1u);
                            mma_tf32_ts_1t(d_tmem, ta + ks * 8, bh, IDESC, 1u);
                        } else {
                            mma_tf32_ts_1t(d_tmem, ta + ks * 8, bh, IDESC, first);
                        }
                    }
                    mma_commit_1t(bar(NS + stage));                              // frees the A stage when the MMAs retire
                    if (!resident) mma_commit_1t(bar(2 * NS + 4 + RB + bslot));  // ... and the weight slot
                    if (kc == kchunks - 1) mma_commit_1t(bar(2 * NS + acc));     // accumulator ready
                    PROF(23);
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                    if (!resident && ++bslot == RB) { bslot = 0; bphase ^= 1; }
                }
                if (++acc == 2) { acc = 0; aphase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == W_LOAD) {
        // =========================== loader (TMA) ===========================
        // One thread feeds both operands.  Weights: bulk copies of the packed images - once per CTA when the n-tile's
        // image fits the ring, otherwise streamed up to RB chunks ahead of the MMAs for every tile.  Gathered operand:
        // per k-chunk one tensor load (Cin >= 32: 32 channels of one filter tap) or two (Cin == 16: 16 channels of two
        // taps) of the tile's pixel rectangle, shifted by the tap offset; pixels outside the image arrive as zeros.
        if (lane == 0) {
            const int nt = blockIdx.x % ntiles;
            const uint64_t tmap = reinterpret_cast<uint64_t>(&a.tmap);
            const bool c16 = a.Cin == 16;
            const uint32_t box_bytes = c16 ? 8192u : 16384u;
            const int KK = a.k * a.k;
            uint32_t b = 0, bphase = 0, rs = 0, rphase = 0;
            if (resident && my_tiles > 0)
                for (int kc = 0; kc < a.kchunks; ++kc) {
                    const float *src = a.wimg + ((size_t)(nt * a.kchunks + kc)) * (PASSES * BN * 32);
                    mbar_expect_tx(bar(2 * NS + 4 + kc), L::B_BYTES);
                    bulk_g2s(sbase + L::B_OFF + kc * L::B_BYTES, src, L::B_BYTES, bar(2 * NS + 4 + kc));
                }
            for (int t = 0; t < my_tiles; ++t) {
                const int m0 = ((blockIdx.x + t * gridDim.x) / ntiles) * TM;
                int n0, h0, w0;
                decode_pix(a, m0, n0, h0, w0);                     // w0 == 0: a tile starts at the beginning of an image row
                const int ch0 = h0 * a.in_stride;
                for (int kc = 0; kc < a.kchunks; ++kc) {
                    if (!resident) {
                        mbar_wait_hint(bar(2 * NS + 4 + RB + b), bphase ^ 1);
                        const float *src = a.wimg + ((size_t)(nt * a.kchunks + kc)) * (PASSES * BN * 32);
                        mbar_expect_tx(bar(2 * NS + 4 + b), L::B_BYTES);
                        bulk_g2s(sbase + L::B_OFF + b * L::B_BYTES, src, L::B_BYTES, bar(2 * NS + 4 + b));
                        if (++b == RB) { b = 0; bphase ^= 1; }
                    }
                    mbar_wait(bar(RBAR + NSLOT + rs), rphase ^ 1);         // landing slot free
                    const uint32_t dst = sbase + L::RAW_OFF + rs * (TM * 128), fb = bar(RBAR + rs);
                    const int *e = a.tab[kc];
                    const bool two = c16 && (kc * 2 + 1) < KK;           // Cin == 16: is the chunk's second tap inside K?
                    mbar_expect_tx(fb, two ? 2 * box_bytes : box_bytes);
                    tma_load_4d(dst, tmap, e[3], e[1], ch0 + e[0], n0, fb);
                    if (two) tma_load_4d(dst + 8192, tmap, e[7], e[5], ch0 + e[4], n0, fb);
                    if (++rs == NSLOT) { rs = 0; rphase ^= 1; }
                }
            }
        }
    } else {
        // =========================== epilogue ===========================
        // two warpgroups: group g owns accumulator g and the CTA's tiles t with t % 2 == g
        constexpr int CB = G::CB, LPR = G::LPR, RPI = G::RPI, NIT = G::NIT, NB = G::NB;
        const int ewarp = warp - W_EPI;              // 0..7
        const int eg = ewarp >> 2, ew = ewarp & 3;   // group, TMEM lane quarter (== warp % 4)
        unsigned char *stg = smem + L::STG_OFF + ewarp * G::WARP_BYTES;
        double *stw = s_stat + ewarp * 2 * BN;
        const int c4 = lane % LPR, rsub = lane / LPR;
        const bool want_stats = (a.out_stats != nullptr) || (a.dz_stats != nullptr);
        const bool fwd = a.wmode == 0;
        // side operand streamed by the epilogue: residual (forward) or x_pre (dgrad mask)
        const float *side = fwd ? a.residual : (a.has_mask ? a.x_pre : nullptr);
        const bool acc_out = !fwd && a.accumulate;
        const uint32_t acc = eg;
        uint32_t aphase = 0;
        for (int t = eg; t < my_tiles; t += 2) {
            const int tile = blockIdx.x + t * gridDim.x;
            const int mt = tile / ntiles;
            const int m = mt * TM + ew * 32 + lane;
            int ob_own = -1;
            if (m < M) {
                int n, ho, wo;
                decode_pix(a, m, n, ho, wo);
                ob_own = ((n * a.Hout + ho * a.out_stride) * a.Wout + wo * a.out_stride) * a.Cn + cta_n0;
            }
            int ob[NIT];                               // element offset of (row i*RPI+rsub, this lane's 4 columns), -1: no row
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
                ob[i] = __shfl_sync(0xffffffffu, ob_own, i * RPI + rsub);
                if (ob[i] >= 0) ob[i] += c4 * 4;
            }
            // side operand: one column batch at a time, issue-all-then-consume (see the producers)
            // two batches in flight (registers): the loads of batch b + 2 are issued as soon as batch b is consumed
            float4 sdq[2][NIT];
            auto side_round = [&](int b) {
#pragma unroll
                for (int i = 0; i < NIT; ++i)
                    sdq[b & 1][i] = (side != nullptr && ob[i] >= 0) ? __ldg(reinterpret_cast<const float4 *>(side + ob[i] + b * CB))
                                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            side_round(0);
            if (NB > 1) side_round(1);
            PROF(30);
            if (lane == 0) mbar_wait_hint(bar(2 * NS + acc), aphase);
            __syncwarp();
            tc_fence_after();
            PROF(31);
            if (DBG(8)) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(2 * NS + 2 + acc));
                aphase ^= 1;
                continue;
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                float v[CB];
#pragma unroll
                for (int qq = 0; qq < CB / 16; ++qq)
                    tmem_ld16_nowait(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BN + b * CB + qq * 16, v + qq * 16);
                tmem_ld_wait();
                PROF(33);
                if (b == NB - 1) {                     // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(2 * NS + 2 + acc));
                }
#pragma unroll
                for (int qq = 0; qq < CB / 4; ++qq)
                    *reinterpret_cast<float4 *>(stg + G::addr(lane, qq)) = make_float4(v[qq * 4], v[qq * 4 + 1], v[qq * 4 + 2], v[qq * 4 + 3]);
                __syncwarp();
                PROF(34);
                const int cl = b * CB + c4 * 4;        // this lane's 4 columns of the n-tile
                float4 cf0, cf1, cf2, cf3;
                if (fwd) {
                    cf0 = *reinterpret_cast<const float4 *>(s_bias + cl);
                } else if (a.has_mask) {
                    cf0 = *reinterpret_cast<const float4 *>(s_msc + cl); cf1 = *reinterpret_cast<const float4 *>(s_msh + cl);
                    cf2 = *reinterpret_cast<const float4 *>(s_mmean + cl); cf3 = *reinterpret_cast<const float4 *>(s_mistd + cl);
                }
                // packed fp32 arithmetic (add / mul / fma .f32x2: the same roundings as the scalar forms, half the issue slots -
                // the epilogue warps share their schedulers with the transform warps that set the k-chunk rate)
                float2 s0a = make_float2(0.f, 0.f), s0b = s0a, s1a = s0a, s1b = s0a;
                const float2 m1 = make_float2(-1.f, -1.f);
#pragma unroll
                for (int i = 0; i < NIT; ++i) {
                    const float4 sd = sdq[b & 1][i];
                    float4 y = *reinterpret_cast<const float4 *>(stg + G::addr(i * RPI + rsub, c4));
                    float2 ya = make_float2(y.x, y.y), yb = make_float2(y.z, y.w);
                    const float2 sda = make_float2(sd.x, sd.y), sdb = make_float2(sd.z, sd.w);
                    const bool valid = ob[i] >= 0;
                    if (fwd) {
                        ya = __fadd2_rn(ya, __fadd2_rn(make_float2(cf0.x, cf0.y), sda));
                        yb = __fadd2_rn(yb, __fadd2_rn(make_float2(cf0.z, cf0.w), sdb));
                        if (valid) {
                            s0a = __fadd2_rn(s0a, ya); s0b = __fadd2_rn(s0b, yb);
                            s1a = __ffma2_rn(ya, ya, s1a); s1b = __ffma2_rn(yb, yb, s1b);
                        }
                    } else {
                        if (acc_out && valid) {
                            const float4 ex = *reinterpret_cast<const float4 *>(a.out + ob[i] + b * CB);
                            ya = __fadd2_rn(ya, make_float2(ex.x, ex.y)); yb = __fadd2_rn(yb, make_float2(ex.z, ex.w));
                        }
                        if (a.has_mask) {
                            const float2 pa = __ffma2_rn(sda, make_float2(cf0.x, cf0.y), make_float2(cf1.x, cf1.y));
                            const float2 pb = __ffma2_rn(sdb, make_float2(cf0.z, cf0.w), make_float2(cf1.z, cf1.w));
                            ya.x = (pa.x > 0.f && valid) ? ya.x : 0.f;
                            ya.y = (pa.y > 0.f && valid) ? ya.y : 0.f;
                            yb.x = (pb.x > 0.f && valid) ? yb.x : 0.f;
                            yb.y = (pb.y > 0.f && valid) ? yb.y : 0.f;
                            s0a = __fadd2_rn(s0a, ya); s0b = __fadd2_rn(s0b, yb);
                            // s1 += (y * (sd - mean)) * inv_std
                            const float2 ta = __ffma2_rn(make_float2(cf2.x, cf2.y), m1, sda), tb = __ffma2_rn(make_float2(cf2.z, cf2.w), m1, sdb);
                            s1a = __ffma2_rn(__fmul2_rn(ya, ta), make_float2(cf3.x, cf3.y), s1a);
                            s1b = __ffma2_rn(__fmul2_rn(yb, tb), make_float2(cf3.z, cf3.w), s1b);
                        }
                    }
                    if (valid) *reinterpret_cast<float4 *>(a.out + ob[i] + b * CB) = make_float4(ya.x, ya.y, yb.x, yb.y);
                }
                const float4 s0 = make_float4(s0a.x, s0a.y, s0b.x, s0b.y), s1 = make_float4(s1a.x, s1a.y, s1b.x, s1b.y);
                if (b + 2 < NB) side_round(b + 2);     // in flight during the statistics and the whole next batch
                PROF(35);
                if (want_stats) {
                    // The 8 lanes with equal c4 (rsub = lane / 4) hold partial sums of the same 8 values (4 columns x
                    // {sum, sum of squares / products}).  Recursive halving: each round a lane keeps half of its values
                    // and hands the other half to its partner, so 4 + 2 + 1 shuffles leave ONE complete sum per lane
                    // (instead of 3 x 8 butterfly shuffles) and every lane does one fp64 read-modify-write.
                    static_assert(LPR == 4, "statistics reduction assumes 4 lanes per row");
                    const bool h2 = (lane & 4) != 0, h3 = (lane & 8) != 0, h4 = (lane & 16) != 0;
                    float k0 = h2 ? s1.x : s0.x, k1 = h2 ? s1.y : s0.y, k2 = h2 ? s1.z : s0.z, k3 = h2 ? s1.w : s0.w;
                    const float g0 = h2 ? s0.x : s1.x, g1 = h2 ? s0.y : s1.y, g2 = h2 ? s0.z : s1.z, g3 = h2 ? s0.w : s1.w;
                    k0 += __shfl_xor_sync(0xffffffffu, g0, 4); k1 += __shfl_xor_sync(0xffffffffu, g1, 4);
                    k2 += __shfl_xor_sync(0xffffffffu, g2, 4); k3 += __shfl_xor_sync(0xffffffffu, g3, 4);
                    float m0 = h3 ? k2 : k0, m1 = h3 ? k3 : k1;
                    const float n0 = h3 ? k0 : k2, n1 = h3 ? k1 : k3;
                    m0 += __shfl_xor_sync(0xffffffffu, n0, 8); m1 += __shfl_xor_sync(0xffffffffu, n1, 8);
                    float t0 = h4 ? m1 : m0;
                    const float u0 = h4 ? m0 : m1;
                    t0 += __shfl_xor_sync(0xffffffffu, u0, 16);
                    // this lane now owns value index (h2, h3, h4) -> kind = h2, column = cl + 2*h3 + h4
                    stw[(h2 ? BN : 0) + cl + (h3 ? 2 : 0) + (h4 ? 1 : 0)] += (double)t0;
                }
                __syncwarp();                          // staging tile is rewritten by the next batch
                PROF(36);
            }
            PROF(32);
            aphase ^= 1;
        }
        if (want_stats) {
            // combine the eight epilogue warps, then ONE fp64 atomic per channel and kind per CTA
            asm volatile("bar.sync 1, 256;" ::: "memory");
            double *st = a.out_stats ? a.out_stats : a.dz_stats;
            for (int idx = ewarp * 32 + lane; idx < 2 * BN; idx += 256) {
                double tsum = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) tsum += s_stat[w * 2 * BN + idx];
                const int kind = idx / BN, col = idx - kind * BN;
                if (!(a.knobs & 4)) atomicAdd(&st[kind * a.Cn + cta_n0 + col], tsum);     // knob 4: timing ablation (wrong statistics)
            }
        }
    }
    PROF(3);
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
    }
    PROF(4);
    if (a.tail) {
        // ---- BatchNorm backward of the tensor this dgrad completes (net/batchnormlayer.py:154-192 through T.grad).
        // The statistics {sum dz, sum dz * xhat} are complete only when EVERY CTA has added its share: grid-wide
        // barrier (the grid is at most one CTA per SM and every CTA is resident or will be without waiting for this
        // one, see launch_tc), then each CTA turns its own dz tiles - still in L2 - into dx.  Same arithmetic, in the
        // same order, as k_bn_bwd_apply (bn.cu): the two paths are bit-identical.
        if (tid == 0) {
            __threadfence();                                   // this CTA's dz tiles and statistics before its arrival
            atomicAdd(a.gbar, 1u);
            unsigned int seen = 0, polls = 0;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.gbar) : "memory");
                if (seen >= gridDim.x) break;
                __nanosleep(64);
            } while (++polls < (1u << 24));                    // ~1 s: never reached unless a CTA of this grid died
            if (seen < gridDim.x) atomicExch(a.gbar + 1, 0xDEADu);   // leave a mark instead of hanging the device
        }
        __syncthreads();
        float *s_k1 = s_msc, *s_mdz = s_msh, *s_mdzx = s_bias;     // reuse: [BN] each (mask scale stays, shift / bias are done)
        if (tid < BN) {
            const int cg = cta_n0 + tid;
            const double cnt = a.mask_bn.count;
            const double sdz = __ldcg(a.dz_stats + cg), sdzx = __ldcg(a.dz_stats + a.Cn + cg);
            s_mdz[tid] = (float)(sdz / cnt);
            s_mdzx[tid] = (float)(sdzx / cnt);
            if ((int)blockIdx.x < ntiles) {                        // one CTA per n-tile owns the parameter gradients
                if (a.dbeta) a.dbeta[cg] += a.pscale * (float)sdz;
                if (a.dgamma) a.dgamma[cg] += a.pscale * (float)sdzx;
            }
        }
        __syncthreads();
        constexpr int Q = BN / 4;                                  // float4 per row of the n-tile
        const float *dzp = a.out, *xp = a.x_pre, *skp = a.tail_skip;
        float *dxp = a.tail_out;
        for (int t = 0; t < my_tiles; ++t) {
            const int mt = (blockIdx.x + t * gridDim.x) / ntiles;
            for (int idx = tid; idx < TM * Q; idx += NTHREADS_CONV) {
                const int r = idx / Q, c = (idx - r * Q) * 4;
                const int m = mt * TM + r;
                if (m >= M) break;
                const size_t o = (size_t)m * a.Cn + cta_n0 + c;    // out_stride == 1: GEMM row m is pixel m
                const float4 g = __ldcg(reinterpret_cast<const float4 *>(dzp + o));
                const float4 xv = __ldg(reinterpret_cast<const float4 *>(xp + o));
                const float gr[4] = {g.x, g.y, g.z, g.w}, xr[4] = {xv.x, xv.y, xv.z, xv.w};
                float rr[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float xh = (xr[j] - s_mmean[c + j]) * s_mistd[c + j];
                    rr[j] = s_k1[c + j] * (gr[j] - s_mdz[c + j] - xh * s_mdzx[c + j]);
                }
                if (skp != nullptr) {
                    const float4 sk = __ldg(reinterpret_cast<const float4 *>(skp + o));
                    rr[0] += sk.x; rr[1] += sk.y; rr[2] += sk.z; rr[3] += sk.w;
                }
                *reinterpret_cast<float4 *>(dxp + o) = make_float4(rr[0], rr[1], rr[2], rr[3]);
            }
        }
    }
}

// ---- weight packing: KC fp32 weights -> per-(n-tile, k-chunk) shared-memory images (hi/lo, swizzled)
struct PackItem {
    const float *w; float *img_fwd; float *img_dgrad;
    int Cin, Cout, k, bn_fwd, bn_dgrad, passes;
};

__global__ void k_pack(const PackItem *__restrict__ items) {
    const PackItem it = items[blockIdx.y];
    const int K = it.k * it.k * it.Cin, Kd = it.k * it.k * it.Cout;
    for (int mode = 0; mode < 2; ++mode) {
        const int BN = mode == 0 ? it.bn_fwd : it.bn_dgrad;
        const int Nout = mode == 0 ? it.Cout : it.Cin;
        const int Kt = mode == 0 ? K : Kd;
        float *img = mode == 0 ? it.img_fwd : it.img_dgrad;
        if (img == nullptr) continue;
        const int kch = (Kt + 31) / 32, nts = Nout / BN;
        const int64_t per = (int64_t)it.passes * BN * 32;
        const int64_t total = (int64_t)nts * kch * BN * 32;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const int e = (int)(i & 3), pch = (int)((i >> 2) & 7), row = (int)((i >> 5) % BN);
            const int64_t blk = i / (BN * 32);
            const int kc = (int)(blk % kch), nt = (int)(blk / kch);
            const int lch = pch ^ (row & 7);
            const int kk = kc * 32 + lch * 4 + e;
            const int n = nt * BN + row;
            float v = 0.f;
            if (kk < Kt) {
                if (mode == 0) {
                    v = it.w[(size_t)kk * it.Cout + n];
                } else {
                    const int tap = kk / it.Cout, o = kk - tap * it.Cout;
                    const int r = tap / it.k, s = tap - r * it.k;
                    const int ftap = (it.k - 1 - r) * it.k + (it.k - 1 - s);
                    v = it.w[(size_t)(ftap * it.Cin + n) * it.Cout + o];
                }
            }
            const uint32_t h = to_tf32(v);
            float *dst = img + blk * per + row * 32 + pch * 4 + e;
            dst[0] = __uint_as_float(h);
            if (it.passes > 1) dst[BN * 32] = __uint_as_float(lo_tf32(v, h));
        }
    }
}

template <int BN, int PASSES>
int launch_tc(const TCArgs &a, cudaStream_t st) {
    using L = SmemLayout<BN, PASSES>;
    static bool done = false;
    if (!done) {
        if (cudaFuncSetAttribute(k_conv_tc<BN, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess)
            return -1;
        done = true;
    }
    int M = a.N * a.Hg * a.Wg;
    int tiles = ((M + TM - 1) / TM) * (a.Cn / BN);
    int ntiles = a.Cn / BN;
    int grid = tiles < 148 ? tiles : 148;
    grid -= grid % ntiles;
    if (launch_pdl(1, k_conv_tc<BN, PASSES>, dim3(grid), dim3(NTHREADS_CONV), L::TOTAL, st, a) != cudaSuccess) return -1;
    return 0;
}

int tc_knobs() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("DPP_TC_KNOBS"); v = e ? atoi(e) : 0; }
    return v;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
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

// Tensor map of the gathered operand.  A tile's 128 GEMM rows must form a rectangle of the gather grid: whole image
// rows (128 % Wg == 0), either box_h = 128 / Wg rows of one image or box_n = 128 / (Hg * Wg) whole images.
// Returns DPP_ENOTSUP for other geometries (the fp32 SIMT kernels take them).
int make_tmap(TCArgs &a) {
    if (a.Cin != 16 && a.Cin % 32 != 0) return DPP_ENOTSUP;
    if (a.Wg > TM || TM % a.Wg != 0) return DPP_ENOTSUP;
    const int rows = TM / a.Wg;
    if (a.Hg >= rows) {
        if (a.Hg % rows != 0) return DPP_ENOTSUP;
        a.box_h = rows; a.box_n = 1;
    } else {
        if (rows % a.Hg != 0) return DPP_ENOTSUP;
        a.box_h = a.Hg; a.box_n = rows / a.Hg;
    }
    if ((reinterpret_cast<uintptr_t>(a.in) & 15) != 0) return DPP_ENOTSUP;
    EncodeTiledFn enc = encode_tiled();
    if (enc == nullptr) return -1;
    const int cbox = a.Cin == 16 ? 16 : 32, s = a.in_stride;
    if (a.Wg * s > 256 || a.box_h * s > 256 || s > 8) return DPP_ENOTSUP;
    cuuint64_t gdim[4] = {(cuuint64_t)a.Cin, (cuuint64_t)a.Win, (cuuint64_t)a.Hin, (cuuint64_t)a.N};
    cuuint64_t gstr[3] = {(cuuint64_t)a.Cin * 4, (cuuint64_t)a.Win * a.Cin * 4, (cuuint64_t)a.Hin * a.Win * a.Cin * 4};
    cuuint32_t box[4] = {(cuuint32_t)cbox, (cuuint32_t)(a.Wg * s), (cuuint32_t)(a.box_h * s), (cuuint32_t)a.box_n};
    cuuint32_t estr[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
    const CUresult r = enc(&a.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(a.in), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, cbox == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

int dispatch_tc(TCArgs &a, int passes, cudaStream_t st) {
    a.knobs = tc_knobs();
    {
        const int rc = make_tmap(a);
        if (rc != 0) return rc;
    }
    const int KK = a.k * a.k;
    for (int kc = 0; kc < a.kchunks; ++kc)
        for (int half = 0; half < 2; ++half) {
            const int k0 = kc * KC + half * 16;
            const int tap = k0 / a.Cin, chan = k0 - tap * a.Cin;
            int *e = a.tab[kc] + half * 4;
            if (tap >= KK) { e[0] = 1 << 28; e[1] = 1 << 28; e[2] = 0; e[3] = 0; continue; }   // beyond K: zero rows
            const int r = tap / a.k, s = tap - r * a.k;
            e[0] = r - a.pad; e[1] = s - a.pad;
            e[2] = ((r - a.pad) * a.Win + (s - a.pad)) * a.Cin + chan;
            e[3] = chan;
        }
    a.wsh = a.hsh = -1;
    if ((a.Wg & (a.Wg - 1)) == 0 && (a.Hg & (a.Hg - 1)) == 0) {
        a.wsh = 0; while ((1 << a.wsh) < a.Wg) ++a.wsh;
        a.hsh = 0; while ((1 << a.hsh) < a.Hg) ++a.hsh;
    }
    const int bn = a.Cn > 128 ? 128 : a.Cn;
#define DPP_TC_CASE(B_)                                                    \
    if (bn == B_) return passes > 1 ? launch_tc<B_, 2>(a, st) : launch_tc<B_, 1>(a, st);
    DPP_TC_CASE(16) DPP_TC_CASE(32) DPP_TC_CASE(64) DPP_TC_CASE(128)
#undef DPP_TC_CASE
    return -1;
}

}  // namespace

int dpp_tc_bn_for(int nout) { return nout > 128 ? 128 : nout; }

extern "C" int dpp_conv_pack_size(int Cin, int Cout, int k, int precision, int64_t *fwd_floats, int64_t *dgrad_floats) {
    DPP_CHECK_ARG(Cin > 0 && Cout > 0 && k > 0 && fwd_floats && dgrad_floats);
    const int passes = precision == 1 ? 2 : 1;
    const int64_t kf = (k * k * Cin + 31) / 32, kd = (k * k * Cout + 31) / 32;
    *fwd_floats = kf * 32 * Cout * passes;
    *dgrad_floats = kd * 32 * Cin * passes;
    return DPP_OK;
}

extern "C" int dpp_conv_pack_all(const void *items_dev, int n_items, void *stream) {
    DPP_CHECK_ARG(items_dev && n_items > 0);
    dim3 grid(32, n_items);
    k_pack<<<grid, 256, 0, S(stream)>>>(reinterpret_cast<const PackItem *>(items_dev));
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

// nout / kin: channels of the produced / gathered tensor (Cout / Cin forward, Cin / Cout backward-data)
static int tc_supported(const dpp_conv_desc *d, int nout, int kin) {
    if (d->precision != 1 && d->precision != 2) return 0;
    if (nout % 16 || nout > 256 || (nout > 128 && nout % 128)) return 0;
    if (kin % 16 || kin > 256) return 0;                               // 64-byte tap segments, BN table of 256 channels
    if ((d->k * d->k * kin + 31) / 32 > 18) return 0;                  // chunk table of the kernel
    // the kernel indexes both tensors with 32-bit element offsets
    if ((int64_t)d->N * d->H * d->W * d->Cin >= (1ll << 31) || (int64_t)d->N * d->Ho * d->Wo * d->Cout >= (1ll << 31)) return 0;
    return 1;
}

int dpp_conv2d_fwd_tc(const dpp_conv_desc *d, const float *x, const dpp_bn_ref *in_bn, const float *w,
                      const float *bias, const float *residual, float *y, double *out_stats, void *stream) {
    (void)w;
    if (!tc_supported(d, d->Cout, d->Cin) || d->wpack_fwd == nullptr) return DPP_ENOTSUP;
    TCArgs a;
    memset(&a, 0, sizeof(a));
    a.in = x; a.wimg = d->wpack_fwd; a.out = y;
    a.N = d->N; a.Hin = d->H; a.Win = d->W; a.Cin = d->Cin;
    a.Hg = d->Ho; a.Wg = d->Wo; a.Hout = d->Ho; a.Wout = d->Wo; a.Cn = d->Cout;
    a.k = d->k; a.pad = d->pad; a.in_stride = d->stride; a.out_stride = 1;
    a.wmode = 0; a.kchunks = (d->k * d->k * d->Cin + 31) / 32;
    if (in_bn) { a.in_bn = *in_bn; a.has_in_bn = 1; }
    a.bias = bias; a.residual = residual; a.out_stats = out_stats;
    {
        const int rc = dispatch_tc(a, d->precision == 1 ? 2 : 1, S(stream));
        if (rc == DPP_ENOTSUP) return DPP_ENOTSUP;
        if (rc != 0) return dpp::fail(DPP_ECUDA, "%s: launch setup failed", __func__);
    }
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

static int dgrad_tc_impl(const dpp_conv_desc *d, const float *dy, float *dx, int accumulate, const dpp_bn_ref *mask_bn,
                         const float *x_pre, double *dz_stats, const float *skip, float *bn_dx, float *dgamma, float *dbeta,
                         float pscale, unsigned int *gbar, void *stream) {
    if (!tc_supported(d, d->Cin, d->Cout) || d->wpack_dgrad == nullptr) return DPP_ENOTSUP;
    TCArgs a;
    memset(&a, 0, sizeof(a));
    a.in = dy; a.wimg = d->wpack_dgrad; a.out = dx;
    a.N = d->N; a.Hin = d->Ho; a.Win = d->Wo; a.Cin = d->Cout;
    a.Hout = d->H; a.Wout = d->W; a.Cn = d->Cin;
    a.k = d->k; a.pad = d->k - 1 - d->pad; a.in_stride = 1;
    if (d->stride == 1) { a.Hg = d->H; a.Wg = d->W; a.out_stride = 1; }
    else { a.Hg = d->Ho; a.Wg = d->Wo; a.out_stride = d->stride; }
    a.wmode = 1; a.kchunks = (d->k * d->k * d->Cout + 31) / 32;
    a.accumulate = accumulate;
    if (mask_bn) { a.mask_bn = *mask_bn; a.has_mask = 1; a.x_pre = x_pre; a.dz_stats = dz_stats; }
    if (bn_dx != nullptr) {
        a.tail = 1; a.tail_skip = skip; a.tail_out = bn_dx; a.dgamma = dgamma; a.dbeta = dbeta; a.pscale = pscale; a.gbar = gbar;
    }
    {
        const int rc = dispatch_tc(a, d->precision == 1 ? 2 : 1, S(stream));
        if (rc == DPP_ENOTSUP) return DPP_ENOTSUP;
        if (rc != 0) return dpp::fail(DPP_ECUDA, "%s: launch setup failed", __func__);
    }
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

int dpp_conv2d_dgrad_tc(const dpp_conv_desc *d, const float *dy, float *dx, int accumulate, const dpp_bn_ref *mask_bn,
                        const float *x_pre, double *dz_stats, void *stream) {
    return dgrad_tc_impl(d, dy, dx, accumulate, mask_bn, x_pre, dz_stats, nullptr, nullptr, nullptr, nullptr, 1.f, nullptr, stream);
}

extern "C" int dpp_conv2d_dgrad_bn_bwd(const dpp_conv_desc *d, const float *dy, const float *w, float *dz, int accumulate,
                                       const dpp_bn_ref *mask_bn, const float *x_pre, double *dz_stats, const float *skip,
                                       float *dx, float *dgamma, float *dbeta, float param_grad_scale, unsigned int *gbar,
                                       void *stream) {
    (void)w;
    DPP_CHECK_ARG(d && dy && dz && mask_bn && x_pre && dz_stats && dx && gbar);
    // the fused tail walks GEMM rows as pixels (stride 1) and needs batch statistics
    if (d->stride != 1 || mask_bn->sums == nullptr) return DPP_ENOTSUP;
    return dgrad_tc_impl(d, dy, dz, accumulate, mask_bn, x_pre, dz_stats, skip, dx, dgamma, dbeta, param_grad_scale, gbar, stream);
}

int dpp_conv2d_wgrad_tc_mn(const dpp_conv_desc *d, const float *x, const dpp_bn_ref *in_bn, const float *dy, float *dw,
                           float *db, void *stream);

int dpp_conv2d_wgrad_tc(const dpp_conv_desc *d, const float *x, const dpp_bn_ref *in_bn, const float *dy, float *dw,
                        float *db, void *stream) {
    if (!tc_supported(d, d->Cout, d->Cin)) return DPP_ENOTSUP;
    return dpp_conv2d_wgrad_tc_mn(d, x, in_bn, dy, dw, db, stream);     // MN-major operand tiles (wgrad_tc_mn.cu)
}

#ifdef DPP_PROFILE
extern "C" int dpp_debug_set_flags(int flags) {
    DPP_CUDA(cudaMemcpyToSymbol(g_dbg, &flags, sizeof(flags)));
    return DPP_OK;
}
extern "C" int dpp_debug_set_prof(void *buf) {
    long long *p = reinterpret_cast<long long *>(buf);
    DPP_CUDA(cudaMemcpyToSymbol(g_prof, &p, sizeof(p)));
    return DPP_OK;
}
#endif
