// Backward-weights on tcgen05 with MN-major operands (the fast wgrad path).
//
//   dW[(r,s,c)][o] += sum_p a[p*stride - pad + (r,s)][c] * dy[p][o],   a = ReLU(BN(x)) recomputed on the fly
//
// The reduction dimension is the PIXEL.  NHWC memory is channel-contiguous, i.e. "M/N-contiguous" for this
// GEMM, so instead of transposing into K-major tiles (conv_tc.cu::k_wgrad_tc, 4-byte scatter stores) the
// tiles are written in the UMMA *MN-major* SWIZZLE_128B canonical layout: one 128-byte row = 32 consecutive
// channels of one pixel, 8 pixel rows per 1024-byte swizzle atom, 32-channel column blocks 4096 bytes apart
// (descriptor: LBO = 4096 B between column blocks, SBO = 1024 B between 8-pixel groups, a_major = b_major =
// MN in the instruction descriptor).  Producers therefore store whole 16-byte pieces, like the forward kernel.
//
// Warp roles: warps 0-15 producers (cp.async into private raw slots, BN+ReLU, TF32 hi/lo split), warp 16 MMA
// issuer; when the pixel loop is done warps 4-7 run the epilogue (TMEM -> red.global.add.v4.f32).
// The pixel loop is bound by the instruction latency of the producer warps (in-kernel timeline, gpu call 10:
// address generation 1300 + transform 850 of a 2650-cycle chunk period with 8 warps), hence 16 warps with a
// mapping that gives every thread ONE pixel per chunk:
//   activations: thread = (pixel j = (t/4) % 32, piece gq = t%4 + 4*(t/128)) and handles pieces gq and gq + 16 of
//     the tile's 32 (tap, 4-channel) pieces: one pixel decode per chunk, tap offsets / BN coefficients in
//     registers; the 4 lanes of a pixel read 64 contiguous bytes; a warp's 16-byte stores cover 8 rows x 4
//     swizzled columns = every bank group 4 times, the minimum 4 wavefronts.  Pieces beyond K are skipped, so a
//     layer with few rows (1x1 16->64: 4 pieces) spreads them over 4 warps instead of loading one warp.
//   dy: thread = (pixel t/16, piece t%16 (+16 for 128 channels)).
#include "tc_common.cuh"
#include <stdlib.h>
#include <vector>

using namespace dpp;
using namespace dpp::tc;

namespace {

#ifdef DPP_PROFILE
// Debug timeline (tools/conv_probe.py): CTA 0 appends (tag, clock64) pairs per role into a global buffer.
__device__ long long *g_prof_wg = nullptr;
#define PROF_DECL(base_) int prof_n_ = (base_); long long *const prof_p_ = blockIdx.x == 0 ? g_prof_wg : nullptr
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

constexpr int TM = 128;
constexpr int WS_R = 16;                    // replicas of the split reduction (see WArgs::ws)
constexpr size_t WS_TILE_FLOATS = (size_t)WS_R * 16 * 128 * 128;   // [16 tiles][R][128 rows][<= 128 columns]
constexpr int NPROD = 16;                   // producer warps
constexpr int W_MMA = NPROD;                // MMA issuer warp
constexpr int NTHREADS = 32 * (NPROD + 1);

struct WArgs {
    const float *x; const float *dy; float *dw; float *db;
    int N, H, W, Cin, Cout, k, stride, pad, Ho, Wo;
    dpp_bn_ref in_bn; int has_in_bn;
    int mtiles, ntiles;
    int splits[8], cps[8];       // per m-tile: pixel splits (CTAs) and 32-pixel chunks per split, sized by the tile's row count
    int lbo16, sbo16, kstep;     // experiment knobs (DPP_MN_LBO / DPP_MN_SBO / DPP_MN_KSTEP), defaults 256 / 32 / 1024
    int knobs;                   // tuning bits (DPP_WG_KNOBS): 1 = L1-allocating activation gathers for k > 1
    int wsh, hsh;                // log2(Wo), log2(Ho) when both are powers of two, else -1
    // Split-pixel reduction without a 148-way atomic pile-up: a CTA adds its tile into replica (split % R) of the
    // library workspace, the last CTA of a tile (arrival counter) folds the R replicas into dW and re-zeroes them.
    float *ws; int *ctr; int R;
};

template <int BN, int PASSES, int NST>
struct Lay {
    static constexpr int BNP = BN < 32 ? 32 : BN;                 // MMA N (padded to one 32-channel block)
    // 4 column blocks of 32 (tap,c) rows x 32 pixels + a 5th block for the "tail" rows of the layer's last, nearly empty
    // m-tile (K = 144 -> 128 + 16, K = 288 -> 256 + 32), which ride along with the m-tile before it (second accumulator)
    static constexpr int A_BLOCKS = BN <= 32 ? 5 : 4;             // (only the 16- and 32-channel layers have such tails)
    static constexpr int A_BYTES = PASSES * A_BLOCKS * 4096;
    static constexpr int B_BYTES = PASSES * (BNP / 32) * 4096;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NDY = BNP / 32;                          // 32-channel column blocks of the dy tile
    static constexpr int NDYH = BNP >= 64 ? BNP / 64 : 1;         // dy pieces (16 B) per producer thread per chunk
    static constexpr int AP = A_BLOCKS > 4 ? 3 : 2;               // activation pieces per transform thread and chunk
    static constexpr int UNITS = (AP + NDYH) | 1;                 // 16-byte units per thread slot, odd: conflict-free LDS.128
    static constexpr int SLOT = UNITS * 16;
    static constexpr int RAW_BYTES = 32 * NPROD * SLOT;
    // raw landing slots = prefetch distance + 1: as many as fit next to the two MMA stages (the loop is bound by
    // the memory latency divided by this distance, not by the producer arithmetic), at most 6
    static constexpr int RD_FIT = (220 * 1024 - NST * STAGE_BYTES) / RAW_BYTES;
    static constexpr int RD = RD_FIT > 8 ? 8 : (RD_FIT < 2 ? 2 : RD_FIT);
    static constexpr int RAW_OFF = NST * STAGE_BYTES;
    static constexpr int BAR_OFF = RAW_OFF + RD * RAW_BYTES;
    static constexpr int COEF_OFF = BAR_OFF + 256;
    static constexpr int TOTAL = COEF_OFF + 2 * 256 * 4 + 1024;
};

// One work item: the pixel chunks [c_begin, c_begin + nchunks) of (m-tile mt, n-tile nt) of one layer.  Called by every
// thread of the CTA with barriers and tensor memory set up, the MMA stages zeroed and the BN coefficients in shared
// memory.  gch0 = chunks this CTA has pushed through the stage ring before (ring position and barrier parities carry
// over from item to item), done_parity = parity of the "accumulator complete" barrier for this item.
// direct: the epilogue reduces straight into dW (grouped launches: few CTAs per tile at any one time).
template <int BN, int PASSES, int NST>
__device__ __forceinline__ void item_run(const WArgs &a, const int mt, const int nt, const int split, const int c_begin,
                                         const int nchunks, const int tail, const bool direct, const uint32_t gch0, const uint32_t done_parity,
                                         unsigned char *smem, const uint32_t sbase, const uint32_t tmem_base,
                                         const float *s_scale, const float *s_shift, float *s_db) {
    using L = Lay<BN, PASSES, NST>;
    constexpr int RD = L::RD, D = RD - 1, BNP = L::BNP, NDY = L::NDY, NDYH = L::NDYH;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    PROF_DECL(lane == 0 ? (warp == 0 ? 0 : warp == W_MMA ? 1000 : warp == 4 ? 2000 : warp == 7 ? 3000 : 4000) : 4000);
    auto bar = [&](int i) { return sbase + L::BAR_OFF + 8 * i; };      // full[s]=s, empty[s]=NST+s, done=2*NST
    const int kd0 = mt * TM, o0 = nt * BN;
    const int Kw = a.k * a.k * a.Cin;
    const int P = a.N * a.Ho * a.Wo;
    if (warp < NPROD) {
        const bool pro = a.has_in_bn != 0, relu = a.in_bn.relu != 0;
        const bool use_ca = (a.knobs & 1) && a.k > 1;
        const bool kspin = (a.knobs & 2) != 0, kskip_p = (a.knobs & 8) != 0;     // experiment knobs (timing ablations)
        const int H = a.H, W = a.W, Cin = a.Cin, Cout = a.Cout, Wo = a.Wo, Ho = a.Ho, stride = a.stride;
        // ---- activation role: pixel j of the chunk, pieces gq and gq + 16 (rows kd0 + piece*4 .. +3 = one tap, 4 channels)
        const int j = (tid >> 2) & 31;
        const int gq = (tid & 3) + 4 * (tid >> 7);
        const uint32_t a_sw = (uint32_t)(j & 3) << 1;
        // (third piece: the tail rows kd0 + 128 .. kd0 + 128 + tail - 1 in column block 4, threads with gq < tail / 4)
        bool g_ok[3];
        int g_dr[3], g_ds[3], g_ch[3];
        uint32_t a_off[3];
        float4 sc[3], sf[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int piece = gq + 16 * k;
            const int kd = kd0 + piece * 4;
            g_ok[k] = k < 2 ? kd < Kw : (L::A_BLOCKS > 4 && gq * 4 < tail);   // rows beyond K are never written: the stage stays zero there
            const int tap = g_ok[k] ? kd / Cin : 0;
            g_ch[k] = g_ok[k] ? kd - tap * Cin : 0;
            g_dr[k] = tap / a.k - a.pad; g_ds[k] = tap % a.k - a.pad;
            a_off[k] = (uint32_t)(piece >> 3) * 4096 + j * 128 + ((((uint32_t)piece & 7) ^ a_sw) << 4);
            sc[k] = make_float4(1.f, 1.f, 1.f, 1.f); sf[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pro && g_ok[k]) { sc[k] = *reinterpret_cast<const float4 *>(s_scale + g_ch[k]); sf[k] = *reinterpret_cast<const float4 *>(s_shift + g_ch[k]); }
        }
        const bool any_a = g_ok[0];                  // piece gq + 16 is valid only if piece gq is
        // ---- dy role: pixel jb of the chunk, piece p16 (+ 16 m) along the channel axis
        const int jb = tid >> 4, p16 = tid & 15;
        const uint32_t b_sw = (uint32_t)(jb & 3) << 1;
        float dbp[NDYH * 4];
#pragma unroll
        for (int i = 0; i < NDYH * 4; ++i) dbp[i] = 0.f;
        uint32_t vbits = 0;       // 3 validity bits per in-flight chunk
        for (int ch = -D; ch < nchunks; ++ch) {
            const int ci = ch + D;
            if (ci < nchunks) {
                const uint32_t slot = sbase + L::RAW_OFF + (ci % RD) * L::RAW_BYTES + tid * L::SLOT;
                const int p0 = (c_begin + ci) * 32;
                uint32_t vm = 0;
                if (any_a) {
                    const int p = p0 + j;
                    int wo, ho, n;
                    if (a.wsh >= 0) { wo = p & (Wo - 1); ho = (p >> a.wsh) & (Ho - 1); n = p >> (a.wsh + a.hsh); }
                    else { wo = p % Wo; ho = (p / Wo) % Ho; n = p / (Wo * Ho); }
                    const float *img = a.x + (size_t)n * H * W * Cin;
                    const int h0 = ho * stride, w0 = wo * stride;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if (!g_ok[k]) continue;
                        const int hi = h0 + g_dr[k], wi = w0 + g_ds[k];
                        const bool v = p < P && (unsigned)hi < (unsigned)H && (unsigned)wi < (unsigned)W;
                        const float *src = v ? img + (hi * W + wi) * Cin + g_ch[k] : a.x;
                        if (use_ca) cp_async16_ca(slot + k * 16, src, v ? 16u : 0u);
                        else cp_async16(slot + k * 16, src, v ? 16u : 0u);
                        vm |= (uint32_t)v << k;
                    }
                }
                {
                    const int pd = p0 + jb;
                    const bool pok = pd < P;
#pragma unroll
                    for (int m = 0; m < NDYH; ++m)
                        if ((p16 + 16 * m) * 4 < BN)
                            cp_async16(slot + (L::AP + m) * 16, pok ? a.dy + (size_t)pd * Cout + o0 + (p16 + 16 * m) * 4 : a.dy, pok ? 16u : 0u);
                }
                const uint32_t sh = 3 * (ci % RD);
                vbits = (vbits & ~(7u << sh)) | (vm << sh);
            }
            cp_async_commit();
            if (ch < 0) continue;
            PROF(10);
            cp_async_wait<D>();
            PROF(11);
            const unsigned char *slot = smem + L::RAW_OFF + (ch % RD) * L::RAW_BYTES + tid * L::SLOT;
            const uint32_t stage = (gch0 + ch) % NST, phase = ((gch0 + ch) / NST) & 1;
            const uint32_t vm = (vbits >> (3 * (ch % RD))) & 7u;
            if (lane == 0) { if (kspin) mbar_spin(bar(NST + stage), phase ^ 1); else mbar_wait(bar(NST + stage), phase ^ 1); }
            __syncwarp();
            PROF(12);
            unsigned char *st = smem + stage * L::STAGE_BYTES;
            if (any_a && !kskip_p) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (!g_ok[k]) continue;
                    float4 xv = *reinterpret_cast<const float4 *>(slot + k * 16);
                    if (pro) {
                        const float2 p0 = __ffma2_rn(make_float2(xv.x, xv.y), make_float2(sc[k].x, sc[k].y), make_float2(sf[k].x, sf[k].y));
                        const float2 p1 = __ffma2_rn(make_float2(xv.z, xv.w), make_float2(sc[k].z, sc[k].w), make_float2(sf[k].z, sf[k].w));
                        xv = make_float4(p0.x, p0.y, p1.x, p1.y);        // packed FMAs: same roundings, half the issue slots
                        if (relu) { xv.x = fmaxf(xv.x, 0.f); xv.y = fmaxf(xv.y, 0.f); xv.z = fmaxf(xv.z, 0.f); xv.w = fmaxf(xv.w, 0.f); }
                    }
                    if (!((vm >> k) & 1u)) xv = make_float4(0.f, 0.f, 0.f, 0.f);
                    unsigned char *dst = st + a_off[k];
                    uint4 h, l;
                    split_tf32x2(xv.x, xv.y, h.x, h.y, l.x, l.y);
                    split_tf32x2(xv.z, xv.w, h.z, h.w, l.z, l.w);
                    *reinterpret_cast<uint4 *>(dst) = h;
                    if (PASSES > 1) *reinterpret_cast<uint4 *>(dst + L::A_BLOCKS * 4096) = l;
                }
            }
#pragma unroll
            for (int m = 0; m < NDYH; ++m) {
                const int pcm = p16 + 16 * m;
                if (pcm * 4 >= BN || kskip_p) continue;
                const float4 bv = *reinterpret_cast<const float4 *>(slot + (L::AP + m) * 16);
                dbp[m * 4] += bv.x; dbp[m * 4 + 1] += bv.y; dbp[m * 4 + 2] += bv.z; dbp[m * 4 + 3] += bv.w;
                unsigned char *dst = st + L::A_BYTES + (pcm >> 3) * 4096 + jb * 128 + ((((uint32_t)pcm & 7) ^ b_sw) << 4);
                uint4 h, l;
                split_tf32x2(bv.x, bv.y, h.x, h.y, l.x, l.y);
                split_tf32x2(bv.z, bv.w, h.z, h.w, l.z, l.w);
                *reinterpret_cast<uint4 *>(dst) = h;
                if (PASSES > 1) *reinterpret_cast<uint4 *>(dst + NDY * 4096) = l;
            }
            PROF(14);
            fence_proxy_async();                 // every writer orders its own stores towards the async proxy ...
            __syncwarp();                        // ... the warp agrees, and one lane publishes the warp's share
            if (lane == 0) mbar_arrive(bar(stage));
            PROF(13);
        }
        PROF(3);
        if (a.db != nullptr && mt == 0) {
            // lanes l and l + 16 hold the same dy piece for the warp's 2 pixels
#pragma unroll
            for (int i = 0; i < NDYH * 4; ++i) {
                float t = dbp[i];
                t += __shfl_xor_sync(0xffffffffu, t, 16);
                const int pcm = p16 + 16 * (i >> 2);
                if (lane < 16 && pcm * 4 < BN) atomicAdd(&s_db[pcm * 4 + (i & 3)], t);     // shared-memory atomic: 16 warps
            }
        }
        if (warp >= 4 && warp < 8 && nchunks > 0) {
            // epilogue (warps 4-7: TMEM lane quarter = warp % 4): row = (tap, c) index, columns = output channels
            const int ew = warp - 4;
            const int kd = kd0 + ew * 32 + lane;
            if (lane == 0) mbar_wait(bar(2 * NST), done_parity);
            __syncwarp();
            tc_fence_after();
            PROF(31);
            if (L::A_BLOCKS > 4 && tail > 0 && ew == 0) {               // the tail rows: second accumulator, TMEM columns [BNP, 2 * BNP), lanes 0 .. tail - 1
#pragma unroll
                for (int cb = 0; cb < BN; cb += 16) {
                    float v[16];
                    tmem_ld16(tmem_base + BNP + cb, v);
                    if (lane < tail) {
                        float *dst = a.dw + (size_t)(kd0 + TM + lane) * a.Cout + o0 + cb;
#pragma unroll
                        for (int t = 0; t < 16; t += 4) red_add_v4(dst + t, v[t], v[t + 1], v[t + 2], v[t + 3]);
                    }
                }
            }
#pragma unroll
            for (int cb = 0; cb < BN; cb += 16) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + cb, v);
                if (kd < Kw) {
                    float *dst = (direct || a.splits[mt] == 1) ? a.dw + (size_t)kd * a.Cout + o0 + cb      // straight into dW
                                                   : a.ws + ((size_t)((mt * a.ntiles + nt) * a.R + split % a.R) * TM + ew * 32 + lane) * BN + cb;
#pragma unroll
                    for (int t = 0; t < 16; t += 4) red_add_v4(dst + t, v[t], v[t + 1], v[t + 2], v[t + 3]);
                }
            }
        }
    } else if (warp == W_MMA) {
        if (nchunks > 0 && elect_one()) {      // ONE thread runs the role: back-to-back UTCHMMA out of uniform registers
            // c_format F32, a/b TF32, a_major = b_major = MN (bits 15, 16), N = BNP, M = 128
            constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                       ((uint32_t)(BNP >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            const uint32_t kstep16 = (uint32_t)a.kstep >> 4;
            for (int ch = 0; ch < nchunks; ++ch) {
                const uint32_t stage = (gch0 + ch) % NST, phase = ((gch0 + ch) / NST) & 1;
                if (a.knobs & 2) mbar_spin(bar(stage), phase); else mbar_wait(bar(stage), phase);
                tc_fence_after();
                PROF(22);
                const uint32_t sa = sbase + stage * L::STAGE_BYTES;
                const uint64_t a0 = make_desc_mn(sa, a.lbo16, a.sbo16), b0 = make_desc_mn(sa + L::A_BYTES, a.lbo16, a.sbo16);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {            // 4 groups of 8 pixels
                    if ((a.knobs & 4) && ks > 0) break;     // knob 4: timing ablation
                    const uint64_t ah = a0 + ks * kstep16, bh = b0 + ks * kstep16;
                    const uint32_t first = (ch == 0 && ks == 0) ? 0u : 1u;
                    if (PASSES > 1) {
                        const uint64_t al = ah + ((L::A_BLOCKS * 4096) >> 4);
                        const uint64_t bl = bh + (((BNP / 32) * 4096) >> 4);
                        mma_tf32_ss_1t(tmem_base, ah, bl, IDESC, first);
                        mma_tf32_ss_1t(tmem_base, al, bh, IDESC, 1u);
                        mma_tf32_ss_1t(tmem_base, ah, bh, IDESC, 1u);
                        if (L::A_BLOCKS > 4 && tail > 0) {     // column block 4 (the tail rows) against the same dy tile; rows >= 32 of this
                                            // accumulator read whatever follows block 4 and are never looked at
                            const uint64_t th = ah + ((4 * 4096) >> 4), tl = al + ((4 * 4096) >> 4);
                            mma_tf32_ss_1t(tmem_base + BNP, th, bl, IDESC, first);
                            mma_tf32_ss_1t(tmem_base + BNP, tl, bh, IDESC, 1u);
                            mma_tf32_ss_1t(tmem_base + BNP, th, bh, IDESC, 1u);
                        }
                    } else {
                        mma_tf32_ss_1t(tmem_base, ah, bh, IDESC, first);
                        if (L::A_BLOCKS > 4 && tail > 0) mma_tf32_ss_1t(tmem_base + BNP, ah + ((4 * 4096) >> 4), bh, IDESC, first);
                    }
                }
                mma_commit_1t(bar(NST + stage));
                if (ch == nchunks - 1) mma_commit_1t(bar(2 * NST));
                PROF(23);
            }
        }
        __syncwarp();
    }
}

template <int BN, int PASSES, int NST>
__global__ void __launch_bounds__(NTHREADS, 1)
k_wgrad_mn(WArgs a) {
    using L = Lay<BN, PASSES, NST>;
    constexpr int BNP = L::BNP;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still in the shared window
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    PROF_DECL(lane == 0 ? (warp == 0 ? 0 : warp == W_MMA ? 1000 : warp == 4 ? 2000 : warp == 7 ? 3000 : 4000) : 4000);
    PROF(1);
    auto bar = [&](int i) { return sbase + L::BAR_OFF + 8 * i; };      // full[s]=s, empty[s]=NST+s, done=2*NST
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::BAR_OFF + 128);
    float *s_scale = reinterpret_cast<float *>(smem + L::COEF_OFF);
    float *s_shift = s_scale + 256;
    constexpr uint32_t TCOLS = BNP <= 32 ? 32 : (BNP <= 64 ? 64 : 128);

    const int nt = blockIdx.x % a.ntiles;
    int split = blockIdx.x / a.ntiles, mt = 0;
    while (mt + 1 < a.mtiles && split >= a.splits[mt]) { split -= a.splits[mt]; ++mt; }
    const int kd0 = mt * TM, o0 = nt * BN;
    const int Kw = a.k * a.k * a.Cin;
    const int P = a.N * a.Ho * a.Wo;
    const int total_chunks = (P + 31) / 32;
    const int c_begin = split * a.cps[mt];
    int c_end = c_begin + a.cps[mt]; if (c_end > total_chunks) c_end = total_chunks;
    const int nchunks = c_end > c_begin ? c_end - c_begin : 0;

    __shared__ float s_db[128];      // this CTA's bias-gradient partial (m-tile 0 only)
    __shared__ int s_last;
    if (tid < 128) s_db[tid] = 0.f;
    pdl_trigger();      // private set-up first (see common.cuh: programmatic dependent launch)
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(bar(s), NPROD); mbar_init(bar(NST + s), 1); }   // full: one arrival per producer warp
        mbar_init(bar(2 * NST), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // zero the MMA stages once: padded dy columns (BN = 16) and rows beyond Kw stay zero
    for (int i = tid; i < NST * L::STAGE_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    pdl_wait();
    if (a.has_in_bn)
        for (int c = tid; c < a.Cin; c += NTHREADS) bn_scale_shift(a.in_bn, c, a.Cin, s_scale[c], s_shift[c]);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    PROF(2);
    item_run<BN, PASSES, NST>(a, mt, nt, split, c_begin, nchunks, 0, false, 0u, 0u, smem, sbase, tmem_base, s_scale, s_shift, s_db);
    PROF(32);
    tc_fence_before();
    __syncthreads();
    PROF(4);
    if (warp == W_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
    }
    // ---- last CTA of this (m-tile, n-tile): fold the replicas into dW (and db)
    const int tile_id = mt * a.ntiles + nt;
    const bool has_db = a.db != nullptr && mt == 0;
    float *const wdb = a.ws + WS_TILE_FLOATS + (size_t)(nt * a.R) * 128;     // [R][128] bias-gradient replicas of n-tile nt
    if (a.splits[mt] == 1) {                          // unsplit tile: nothing to fold
        if (has_db && tid < BN) a.db[o0 + tid] += s_db[tid];
        return;
    }
    if (has_db && tid < BN) atomicAdd(wdb + (split % a.R) * 128 + tid, s_db[tid]);   // s_db is complete: __syncthreads above
    __threadfence();                                  // this CTA's reductions are ordered before its arrival
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&a.ctr[tile_id], 1) == a.splits[mt] - 1);
    __syncthreads();
    if (s_last) {
        __threadfence();
        const int rows = Kw - kd0 < TM ? Kw - kd0 : TM;
        float *base = a.ws + (size_t)tile_id * a.R * TM * BN;
        for (int e = tid; e < rows * (BN / 4); e += NTHREADS) {
            const int row = e / (BN / 4), c4 = e - row * (BN / 4);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < a.R; ++r) {
                float4 *p = reinterpret_cast<float4 *>(base + ((size_t)r * TM + row) * BN) + c4;
                const float4 v = __ldcg(p);           // written by other SMs' L2 reductions
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                __stcg(p, make_float4(0.f, 0.f, 0.f, 0.f));   // leave the workspace clean for the next launch
            }
            float4 *d = reinterpret_cast<float4 *>(a.dw + (size_t)(kd0 + row) * a.Cout + o0) + c4;
            float4 o = *d;
            o.x += acc.x; o.y += acc.y; o.z += acc.z; o.w += acc.w;
            *d = o;
        }
        if (has_db && tid < BN) {
            float acc = 0.f;
            for (int r = 0; r < a.R; ++r) { acc += __ldcg(wdb + r * 128 + tid); __stcg(wdb + r * 128 + tid, 0.f); }
            a.db[o0 + tid] += acc;
        }
        if (tid == 0) a.ctr[tile_id] = 0;
    }
}

// -------------------------------------------------------------------------------------------------------------------
// Grouped launch: the backward-weights GEMMs of MANY layers in one persistent kernel.  A step has 63 of them, most far
// too small to fill the device for longer than their own fixed cost (launch, set-up, pipeline fill, split reduction:
// ~9 us each, a third of their total time); their results are needed only by the optimiser.  The work of all layers with
// the same n-tile width is cut into items of `nchunks` pixel chunks of one (layer, m-tile, n-tile); CTAs (one per SM)
// take items from a list - the first statically, the rest through an atomic counter - and reduce each item's tile
// straight into dW (the list interleaves the tiles, so only a handful of CTAs work on the same tile at any time).
// -------------------------------------------------------------------------------------------------------------------
struct WItem { int layer, mt, nt, c_begin, nchunks, tail, pad1, pad2; };   // tail: rows of the layer's last m-tile merged into this one
struct WGroupArgs { const WArgs *layers; const WItem *items; int n_items; unsigned int *sched; };

template <int BN, int PASSES, int NST>
__global__ void __launch_bounds__(NTHREADS, 1)
k_wgrad_group(const WGroupArgs g) {
    using L = Lay<BN, PASSES, NST>;
    constexpr int BNP = L::BNP;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5;
    auto bar = [&](int i) { return sbase + L::BAR_OFF + 8 * i; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::BAR_OFF + 128);
    float *s_scale = reinterpret_cast<float *>(smem + L::COEF_OFF);
    float *s_shift = s_scale + 256;
    constexpr uint32_t TCOLS = 2 * (BNP <= 32 ? 32 : (BNP <= 64 ? 64 : 128));     // two accumulators (m-tile + merged tail rows)
    __shared__ WArgs s_a;
    __shared__ float s_db[128];
    __shared__ int s_next;

    pdl_trigger();
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(bar(s), NPROD); mbar_init(bar(NST + s), 1); }
        mbar_init(bar(2 * NST), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    uint32_t gch = 0, done_parity = 0;
    int item = blockIdx.x;
    while (item < g.n_items) {
        const WItem it = g.items[item];
        if (tid == 0) s_next = (int)(atomicAdd(g.sched, 1u) + gridDim.x);          // the item after this one
        {   // the layer's arguments -> shared memory
            const uint32_t *src = reinterpret_cast<const uint32_t *>(g.layers + it.layer);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&s_a);
            for (int i = tid; i < (int)(sizeof(WArgs) / 4); i += NTHREADS) dst[i] = src[i];
        }
        if (tid < 128) s_db[tid] = 0.f;
        // zero the MMA stages: rows beyond K and padded dy columns of THIS layer must be zero (the previous item's MMAs
        // have retired: its epilogue waited for them before the barrier at the end of the loop)
        for (int i = tid; i < NST * L::STAGE_BYTES / 16; i += NTHREADS) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        if (s_a.has_in_bn)
            for (int c = tid; c < s_a.Cin; c += NTHREADS) bn_scale_shift(s_a.in_bn, c, s_a.Cin, s_scale[c], s_shift[c]);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        item_run<BN, PASSES, NST>(s_a, it.mt, it.nt, 0, it.c_begin, it.nchunks, it.tail, true, gch, done_parity, smem, sbase,
                                  tmem_base, s_scale, s_shift, s_db);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (s_a.db != nullptr && it.mt == 0 && tid < BN) atomicAdd(s_a.db + it.nt * BN + tid, s_db[tid]);
        gch += (uint32_t)it.nchunks;
        done_parity ^= 1u;
        item = s_next;
        __syncthreads();                // s_next / s_a / s_db are rewritten by the next round
    }
    if (warp == W_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
    }
}

// library-owned workspace of the split reduction (16 replicas x up to 16 tiles of 128 x 128 floats) + arrival counters
constexpr size_t WS_FLOATS = WS_TILE_FLOATS + (size_t)2 * WS_R * 128;
float *g_ws = nullptr;
int *g_ctr = nullptr;

int ws_init() {
    if (g_ws != nullptr) return 0;
    float *w = nullptr; int *c = nullptr;
    if (cudaMalloc(&w, WS_FLOATS * sizeof(float)) != cudaSuccess) return -1;
    if (cudaMalloc(&c, 64 * sizeof(int)) != cudaSuccess) { cudaFree(w); return -1; }
    if (cudaMemset(w, 0, WS_FLOATS * sizeof(float)) != cudaSuccess || cudaMemset(c, 0, 64 * sizeof(int)) != cudaSuccess ||
        cudaDeviceSynchronize() != cudaSuccess) { cudaFree(w); cudaFree(c); return -1; }
    g_ws = w; g_ctr = c;
    return 0;
}

template <int BN, int PASSES, int NST>
int launch(WArgs &a, cudaStream_t st) {
    using L = Lay<BN, PASSES, NST>;
    static bool done = false;
    if (!done) {
        if (cudaFuncSetAttribute(k_wgrad_mn<BN, PASSES, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess) return -1;
        done = true;
    }
    const int Kw = a.k * a.k * a.Cin;
    a.mtiles = (Kw + TM - 1) / TM;
    a.ntiles = a.Cout / BN;
    if (a.mtiles > 8) return -1;
    const int P = a.N * a.Ho * a.Wo;
    const int total_chunks = (P + 31) / 32;
    // CTA budget per n-tile, shared between the m-tiles in proportion to their producer work per pixel chunk:
    // A pieces (valid rows / 4) + dy pieces (BN / 4) + a fixed synchronisation cost.  The last m-tile of a 3x3
    // layer often holds only a few rows (K = 144 -> 128 + 16) and gets correspondingly fewer CTAs.
    int cost[8], total_cost = 0, grid = 0;
    for (int t = 0; t < a.mtiles; ++t) {
        const int rows = Kw - t * TM < TM ? Kw - t * TM : TM;
        cost[t] = rows / 4 + BN / 4 + 64;       // the fixed term is the per-chunk pipeline latency, which dominates
        total_cost += cost[t];
    }
    int budget = 148 / a.ntiles; if (budget < a.mtiles) budget = a.mtiles;
    int used = 0;
    for (int t = 0; t < a.mtiles; ++t) {
        int sp = budget * cost[t] / total_cost; if (sp < 1) sp = 1;
        if (sp > total_chunks) sp = total_chunks;
        a.splits[t] = sp; used += sp;
    }
    for (int guard = 0; used < budget && guard < 256; ++guard) {      // hand out the remainder to the most loaded tiles
        int best = -1;
        for (int t = 0; t < a.mtiles; ++t)
            if (a.splits[t] < total_chunks && (best < 0 || (int64_t)cost[t] * a.splits[best] > (int64_t)cost[best] * a.splits[t])) best = t;
        if (best < 0) break;
        ++a.splits[best]; ++used;
    }
    for (int t = 0; t < a.mtiles; ++t) {
        a.cps[t] = (total_chunks + a.splits[t] - 1) / a.splits[t];
        a.splits[t] = (total_chunks + a.cps[t] - 1) / a.cps[t];
        grid += a.splits[t] * a.ntiles;
    }
    if (ws_init() != 0 || a.mtiles * a.ntiles > 16) return -1;
    a.ws = g_ws; a.ctr = g_ctr;
    // replicas: enough to keep the same-address reduction chain at ~8 CTAs, no more (the last CTA folds R tiles)
    int smin = a.splits[0], smax = a.splits[0];
    for (int t = 1; t < a.mtiles; ++t) { if (a.splits[t] < smin) smin = a.splits[t]; if (a.splits[t] > smax) smax = a.splits[t]; }
    a.R = (smax + 7) / 8;
    if (a.R > WS_R) a.R = WS_R;
    if (a.R > smin) a.R = smin;
    if (a.R < 1) a.R = 1;
    if (launch_pdl(2, k_wgrad_mn<BN, PASSES, NST>, dim3(grid), dim3(NTHREADS), L::TOTAL, st, a) != cudaSuccess) return -1;
    return 0;
}

}  // namespace

static void fill_args(WArgs &a, const dpp_conv_desc *d, const float *x, const dpp_bn_ref *in_bn, const float *dy, float *dw,
                      float *db) {
    memset(&a, 0, sizeof(a));
    a.x = x; a.dy = dy; a.dw = dw; a.db = db;
    a.N = d->N; a.H = d->H; a.W = d->W; a.Cin = d->Cin; a.Cout = d->Cout;
    a.k = d->k; a.stride = d->stride; a.pad = d->pad; a.Ho = d->Ho; a.Wo = d->Wo;
    if (in_bn) { a.in_bn = *in_bn; a.has_in_bn = 1; }
    { const char *e; a.lbo16 = (e = getenv("DPP_MN_LBO")) ? atoi(e) : 256; a.sbo16 = (e = getenv("DPP_MN_SBO")) ? atoi(e) : 32;
      a.kstep = (e = getenv("DPP_MN_KSTEP")) ? atoi(e) : 1024;
      a.knobs = (e = getenv("DPP_WG_KNOBS")) ? atoi(e) : 0; }
    a.wsh = a.hsh = -1;
    if ((a.Wo & (a.Wo - 1)) == 0 && (a.Ho & (a.Ho - 1)) == 0) {
        a.wsh = 0; while ((1 << a.wsh) < a.Wo) ++a.wsh;
        a.hsh = 0; while ((1 << a.hsh) < a.Ho) ++a.hsh;
    }
}

static int wgrad_mn_supported(const dpp_conv_desc *d) {
    if (d->precision != 1 && d->precision != 2) return 0;
    if (d->Cout % 16 || d->Cout > 256 || (d->Cout > 128 && d->Cout % 128)) return 0;
    if (d->Cin % 4 || d->Cin > 256 || (d->k * d->k * d->Cin + TM - 1) / TM > 8) return 0;
    return 1;
}

int dpp_conv2d_wgrad_tc_mn(const dpp_conv_desc *d, const float *x, const dpp_bn_ref *in_bn, const float *dy, float *dw,
                           float *db, void *stream) {
    if (!wgrad_mn_supported(d)) return DPP_ENOTSUP;
    WArgs a;
    fill_args(a, d, x, in_bn, dy, dw, db);
    const int bn = d->Cout > 128 ? 128 : d->Cout;
    const bool p3 = d->precision == 1;
    // MMA stages: two by default.  A third stage (DPP_WG_NST=3, BN <= 64) was measured on B200 and changes nothing
    // (A 76.9 vs 76.5 us, B 56.6 vs 56.2 us): the loop is not bound by the stage hand-off but by the load path.
    static int nst_env = -1;
    if (nst_env < 0) { const char *e = getenv("DPP_WG_NST"); nst_env = e ? atoi(e) : 2; }
    const bool three = nst_env >= 3 && bn <= 64;
    int rc = -1;
#define DPP_WG_CASE(B_)                                                                                   \
    if (bn == B_) rc = p3 ? (three ? launch<B_, 2, 3>(a, S(stream)) : launch<B_, 2, 2>(a, S(stream)))      \
                          : (three ? launch<B_, 1, 3>(a, S(stream)) : launch<B_, 1, 2>(a, S(stream)));
    DPP_WG_CASE(16) DPP_WG_CASE(32) DPP_WG_CASE(64)
#undef DPP_WG_CASE
    if (bn == 128) rc = p3 ? launch<128, 2, 2>(a, S(stream)) : launch<128, 1, 2>(a, S(stream));
    if (rc != 0) return dpp::fail(DPP_ECUDA, "%s: launch setup failed", __func__);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

#ifdef DPP_PROFILE
extern "C" int dpp_debug_set_prof_wg(void *buf) {
    long long *p = reinterpret_cast<long long *>(buf);
    DPP_CUDA(cudaMemcpyToSymbol(g_prof_wg, &p, sizeof(p)));
    return DPP_OK;
}
#endif

// ---- grouped backward-weights ------------------------------------------------------------------------------------------
namespace dpp {
// wgrad_simt3.cu: the 3x3 Cin = Cout = 16 / 32 layers run as an fp32 FMA kernel (faster than tcgen05 at N = 16 / 32)
struct W3Layer {
    const float *x; const float *dy; float *dw; float *db;
    dpp_bn_ref in_bn; int has_in_bn;
    int N, H, W, C;
    int rblocks, item0;
    unsigned wp_magic, w_magic;
    int cout;
};
int wgrad3_kind(const dpp_conv_desc *d);        // 1 = 3x3 16 / 32 channels, 2 = 1x1 64 -> 16 / 16 -> 64, 0 = stays on tcgen05
int wgrad3_create(const std::vector<W3Layer> &layers, void **handle_out);
int wgrad3_launches(void *handle);
int wgrad3_run(void *handle, cudaStream_t st);
void wgrad3_destroy(void *handle);
}

namespace {

struct WGroupLaunch { int bn; WGroupArgs args; int grid; };
struct WGroup {
    std::vector<WGroupLaunch> launches;
    std::vector<void *> allocs;
    int passes;
    void *w3 = nullptr;       // handle of the fp32 3x3 launches (wgrad_simt3.cu)
};

// A layer whose last m-tile holds at most 32 rows (K = 144 -> 128 + 16, K = 288 -> 256 + 32) does not get a pass over all
// pixels for those rows alone: they ride with the m-tile before (WItem::tail, second accumulator in item_run)
void merged_tail(int bn, int Kw, int &mtiles, int &tail) {
    tail = 0;
    const int last = Kw - (mtiles - 1) * TM;
    if (bn <= 32 && mtiles >= 2 && last <= 32 && last % 4 == 0) { tail = last; --mtiles; }
}

template <int BN, int PASSES>
int launch_group(const WGroupLaunch &l, cudaStream_t st) {
    using L = Lay<BN, PASSES, 2>;
    static bool done = false;
    if (!done) {
        if (cudaFuncSetAttribute(k_wgrad_group<BN, PASSES, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess) return -1;
        done = true;
    }
    if (cudaMemsetAsync(l.args.sched, 0, sizeof(unsigned int), st) != cudaSuccess) return -1;
    if (launch_pdl(2, k_wgrad_group<BN, PASSES, 2>, dim3(l.grid), dim3(NTHREADS), L::TOTAL, st, l.args) != cudaSuccess) return -1;
    return 0;
}

}  // namespace

extern "C" int dpp_wgrad_group_create(const dpp_wgrad_layer *layers, int n_layers, void **handle_out) {
    DPP_CHECK_ARG(layers && n_layers > 0 && handle_out);
    for (int i = 0; i < n_layers; ++i) {
        DPP_CHECK_ARG(layers[i].x && layers[i].dy && layers[i].dw);
        if (!wgrad_mn_supported(&layers[i].d)) return DPP_ENOTSUP;
        DPP_CHECK_ARG(layers[i].d.precision == layers[0].d.precision);
    }
    WGroup *grp = new WGroup();
    grp->passes = layers[0].d.precision == 1 ? 2 : 1;
    {
        std::vector<dpp::W3Layer> l3;
        for (int i = 0; i < n_layers; ++i)
            if (const int kind = dpp::wgrad3_kind(&layers[i].d)) {
                dpp::W3Layer w;
                memset(&w, 0, sizeof(w));
                w.x = layers[i].x; w.dy = layers[i].dy; w.dw = layers[i].dw; w.db = layers[i].db;
                w.has_in_bn = layers[i].has_in_bn;
                if (w.has_in_bn) w.in_bn = layers[i].in_bn;
                w.N = layers[i].d.N; w.H = layers[i].d.H; w.W = layers[i].d.W; w.C = layers[i].d.Cin;
                w.cout = layers[i].d.Cout;
                if (kind == 2) { w.H = layers[i].d.N * layers[i].d.H * layers[i].d.W; w.W = 1; w.N = 1; }   // a flat pixel range
                l3.push_back(w);
            }
        if (!l3.empty() && dpp::wgrad3_create(l3, &grp->w3) != 0) {
            delete grp;
            return dpp::fail(DPP_ECUDA, "%s: set-up of the fp32 3x3 launches failed", __func__);
        }
    }
    const int widths[4] = {16, 32, 64, 128};
    for (int wi = 0; wi < 4; ++wi) {
        const int bn = widths[wi];
        std::vector<WArgs> la;
        for (int i = 0; i < n_layers; ++i) {
            const int lbn = layers[i].d.Cout > 128 ? 128 : layers[i].d.Cout;
            if (lbn != bn || dpp::wgrad3_kind(&layers[i].d) != 0) continue;
            WArgs a;
            fill_args(a, &layers[i].d, layers[i].x, layers[i].has_in_bn ? &layers[i].in_bn : nullptr, layers[i].dy, layers[i].dw,
                      layers[i].db);
            la.push_back(a);
        }
        if (la.empty()) continue;
        // items: every (layer, m-tile, n-tile) cut into runs of `cpi` pixel chunks; cpi such that a CTA sees ~6 items
        long long total = 0;
        for (const WArgs &a : la) {
            const int Kw = a.k * a.k * a.Cin;
            int mtiles = (Kw + TM - 1) / TM, tail = 0;
            merged_tail(bn, Kw, mtiles, tail);
            total += (long long)mtiles * (a.Cout / bn) * ((a.N * a.Ho * a.Wo + 31) / 32);
        }
        long long cpi = (total + 148 * 6 - 1) / (148 * 6);
        if (cpi < 8) cpi = 8;
        if (cpi > 64) cpi = 64;
        // list order: round s of every tile, then round s + 1, ... so that neighbouring items (which run at the same
        // time) belong to different tiles and do not pile their reductions onto the same addresses
        std::vector<WItem> items;
        for (int s = 0;; ++s) {
            bool any = false;
            for (size_t li = 0; li < la.size(); ++li) {
                const WArgs &a = la[li];
                const int Kw = a.k * a.k * a.Cin, ntiles = a.Cout / bn;
                int mtiles = (Kw + TM - 1) / TM, tail = 0;
                merged_tail(bn, Kw, mtiles, tail);
                const int chunks = (a.N * a.Ho * a.Wo + 31) / 32;
                const int c0 = (int)(s * cpi);
                if (c0 >= chunks) continue;
                any = true;
                const int n = chunks - c0 < cpi ? chunks - c0 : (int)cpi;
                for (int mt = 0; mt < mtiles; ++mt)
                    for (int nt = 0; nt < ntiles; ++nt) {
                        WItem it;
                        memset(&it, 0, sizeof(it));
                        it.layer = (int)li; it.mt = mt; it.nt = nt; it.c_begin = c0; it.nchunks = n;
                        it.tail = mt == mtiles - 1 ? tail : 0;
                        items.push_back(it);
                    }
            }
            if (!any) break;
        }
        WGroupLaunch l;
        l.bn = bn;
        void *dl = nullptr, *di = nullptr, *ds = nullptr;
        if (cudaMalloc(&dl, la.size() * sizeof(WArgs)) != cudaSuccess || cudaMalloc(&di, items.size() * sizeof(WItem)) != cudaSuccess ||
            cudaMalloc(&ds, 64) != cudaSuccess ||
            cudaMemcpy(dl, la.data(), la.size() * sizeof(WArgs), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(di, items.data(), items.size() * sizeof(WItem), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemset(ds, 0, 64) != cudaSuccess) {
            delete grp;
            return dpp::fail(DPP_ECUDA, "%s: device table allocation failed", __func__);
        }
        grp->allocs.push_back(dl); grp->allocs.push_back(di); grp->allocs.push_back(ds);
        l.args.layers = reinterpret_cast<const WArgs *>(dl);
        l.args.items = reinterpret_cast<const WItem *>(di);
        l.args.n_items = (int)items.size();
        l.args.sched = reinterpret_cast<unsigned int *>(ds);
        l.grid = l.args.n_items < 148 ? l.args.n_items : 148;
        grp->launches.push_back(l);
    }
    *handle_out = grp;
    return DPP_OK;
}

extern "C" int dpp_wgrad_group_run(void *handle, void *stream) {
    DPP_CHECK_ARG(handle);
    WGroup *grp = reinterpret_cast<WGroup *>(handle);
    if (grp->w3 != nullptr && dpp::wgrad3_run(grp->w3, S(stream)) != 0) return dpp::fail(DPP_ECUDA, "%s: fp32 3x3 launch failed", __func__);
    for (const WGroupLaunch &l : grp->launches) {
        int rc = -1;
        const bool p3 = grp->passes == 2;
        if (l.bn == 16) rc = p3 ? launch_group<16, 2>(l, S(stream)) : launch_group<16, 1>(l, S(stream));
        else if (l.bn == 32) rc = p3 ? launch_group<32, 2>(l, S(stream)) : launch_group<32, 1>(l, S(stream));
        else if (l.bn == 64) rc = p3 ? launch_group<64, 2>(l, S(stream)) : launch_group<64, 1>(l, S(stream));
        else if (l.bn == 128) rc = p3 ? launch_group<128, 2>(l, S(stream)) : launch_group<128, 1>(l, S(stream));
        if (rc != 0) return dpp::fail(DPP_ECUDA, "%s: launch setup failed", __func__);
        DPP_LAUNCH_CHECK();
    }
    return DPP_OK;
}

extern "C" int dpp_wgrad_group_launches(void *handle) {
    if (!handle) return 0;
    WGroup *grp = reinterpret_cast<WGroup *>(handle);
    return (int)grp->launches.size() + (grp->w3 ? dpp::wgrad3_launches(grp->w3) : 0);
}

extern "C" int dpp_wgrad_group_destroy(void *handle) {
    if (handle) {
        WGroup *grp = reinterpret_cast<WGroup *>(handle);
        for (void *p : grp->allocs) cudaFree(p);
        if (grp->w3) dpp::wgrad3_destroy(grp->w3);
        delete grp;
    }
    return DPP_OK;
}

// Allocates the library-owned workspace of the backward-weights split reduction (idempotent).  Must be called
// outside CUDA-graph capture (Engine does it at construction); dpp_conv2d_wgrad allocates lazily otherwise.
extern "C" int dpp_wgrad_workspace_init(void) {
    if (ws_init() != 0) return dpp::fail(DPP_ECUDA, "%s: workspace allocation failed", __func__);
    return DPP_OK;
}
