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
// Warp roles / pipeline are those of k_wgrad_tc: warps 0-3 producers (cp.async into private raw slots,
// BN+ReLU, TF32 hi/lo split), warp 8 lane 0 MMA issuer, warps 4-7 epilogue (TMEM -> red.global.add.v4.f32).
#include "common.cuh"
#include <stdlib.h>

using namespace dpp;

namespace {

constexpr int TM = 128;
constexpr int NTHREADS = 288;
constexpr int NST = 2;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// TF32 operand = fp32 with the low 13 mantissa bits cleared (truncation).  hi = trunc(x), lo = trunc(x - hi):
// x - hi is exact in fp32, so hi + lo reproduces x to 2^-21 relative - the 3xTF32 split in 3 ALU ops per value
// (cvt.rna.tf32.f32 expands to a ~10-instruction sequence on sm_100a and dominated the producer loop).
__device__ __forceinline__ uint32_t to_tf32(float x) { return __float_as_uint(x) & 0xFFFFE000u; }
// MN-major SWIZZLE_128B descriptor: LBO = 4096 B (next 32-channel block), SBO = 1024 B (next 8 pixels)
// MN-major tf32 operands accept only the SWIZZLE_128B_BASE32B layout (type 1): atoms of 4 K-rows x 128 B,
// 32-byte chunk index XOR (K-row & 3); LBO = stride between 32-element MN blocks, SBO = between 4-row K atoms.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo16 = 256, uint32_t sbo16 = 32, uint32_t lt = 1) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)lbo16 << 16) | ((uint64_t)sbo16 << 32) | (1ull << 46) | ((uint64_t)lt << 61);
}
// executed by a CONVERGED warp: elect.sync inside the asm lets ptxas emit a bare UTCHMMA (a lane-0 branch
// around tcgen05.mma costs an ELECT/BRA.U.ANY loop of ~50 stall cycles per instruction)
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b32 r;\n\t"
        "elect.sync r|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t.reg .b32 r;\n\t"
        "elect.sync r|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void red_add_v4(float *dst, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct WArgs {
    const float *x; const float *dy; float *dw; float *db;
    int N, H, W, Cin, Cout, k, stride, pad, Ho, Wo;
    dpp_bn_ref in_bn; int has_in_bn;
    int mtiles, ntiles, splits, chunks_per_split;
    int lbo16, sbo16, kstep;     // experiment knobs (DPP_MN_LBO / DPP_MN_SBO / DPP_MN_KSTEP), defaults 256 / 32 / 1024
    int knobs;                   // tuning bits (DPP_WG_KNOBS): 1 = L1-allocating activation gathers for k > 1
};

template <int BN, int PASSES>
struct Lay {
    static constexpr int BNP = BN < 32 ? 32 : BN;                 // MMA N (padded to one 32-channel block)
    static constexpr int A_BYTES = PASSES * 4 * 4096;             // 4 column blocks of 32 (tap,c) rows x 32 pixels
    static constexpr int B_BYTES = PASSES * (BNP / 32) * 4096;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NPB = BN / 16;                           // dy pieces (16 B) per thread per chunk
    static constexpr int SLOT = (8 + NPB) * 16 + 16;
    static constexpr int RAW_BYTES = 128 * SLOT;
    static constexpr int RD = (NST * STAGE_BYTES + 3 * RAW_BYTES + 4096 <= 225 * 1024) ? 3 : 2;
    static constexpr int RAW_OFF = NST * STAGE_BYTES;
    static constexpr int BAR_OFF = RAW_OFF + RD * RAW_BYTES;
    static constexpr int COEF_OFF = BAR_OFF + 256;
    static constexpr int TOTAL = COEF_OFF + 2 * 256 * 4 + 1024;
};

template <int BN, int PASSES>
__global__ void __launch_bounds__(NTHREADS, 1)
k_wgrad_mn(WArgs a) {
    using L = Lay<BN, PASSES>;
    constexpr int RD = L::RD, D = RD - 1, BNP = L::BNP, NPB = L::NPB;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-byte aligned, still in the shared window
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto bar = [&](int i) { return sbase + L::BAR_OFF + 8 * i; };      // full[s]=s, empty[s]=NST+s, done=2*NST
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::BAR_OFF + 128);
    float *s_scale = reinterpret_cast<float *>(smem + L::COEF_OFF);
    float *s_shift = s_scale + 256;
    constexpr uint32_t TCOLS = BNP <= 32 ? 32 : (BNP <= 64 ? 64 : 128);

    const int tile = blockIdx.x % (a.mtiles * a.ntiles), split = blockIdx.x / (a.mtiles * a.ntiles);
    const int mt = tile / a.ntiles, nt = tile % a.ntiles;
    const int kd0 = mt * TM, o0 = nt * BN;
    const int Kw = a.k * a.k * a.Cin;
    const int P = a.N * a.Ho * a.Wo;
    const int total_chunks = (P + 31) / 32;
    const int c_begin = split * a.chunks_per_split;
    int c_end = c_begin + a.chunks_per_split; if (c_end > total_chunks) c_end = total_chunks;
    const int nchunks = c_end > c_begin ? c_end - c_begin : 0;

    pdl_trigger();      // private set-up first (see common.cuh: programmatic dependent launch)
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(bar(s), 128); mbar_init(bar(NST + s), 1); }
        mbar_init(bar(2 * NST), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {
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

    if (warp < 4) {
        // producers: thread = (pixel j of the 32-pixel chunk, column quarter q)
        const int j = tid & 31, q = tid >> 5;
        const bool pro = a.has_in_bn != 0, relu = a.in_bn.relu != 0;
        const bool use_ca = (a.knobs & 1) && a.k > 1;
        const int H = a.H, W = a.W, Cin = a.Cin, Cout = a.Cout, Wo = a.Wo, Ho = a.Ho, stride = a.stride;
        // piece g of this thread covers rows kd0 + q*32 + g*4 .. +3 = one tap, 4 channels
        int g_dr[8], g_ds[8], g_ch[8];
        bool g_ok[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const int kd = kd0 + q * 32 + g * 4;
            g_ok[g] = kd < Kw;
            const int tap = g_ok[g] ? kd / Cin : 0;
            g_ch[g] = g_ok[g] ? kd - tap * Cin : 0;
            g_dr[g] = tap / a.k - a.pad; g_ds[g] = tap % a.k - a.pad;
        }
        float dbp[NPB * 4];
#pragma unroll
        for (int i = 0; i < NPB * 4; ++i) dbp[i] = 0.f;
        // within-stage byte offset of (block, pixel row j, 16-byte piece c): block*4096 + j*128 + ((c ^ ((j&3)<<1))<<4)
        const uint32_t rowoff = j * 128;
        const uint32_t sw = (j & 3) << 1;
        // pixel cursor of the issue side, advanced by 32 pixels per chunk without divisions
        int p_i = c_begin * 32 + j;
        int wo_i = p_i % Wo, ho_i = (p_i / Wo) % Ho, n_i = p_i / (Wo * Ho);
        uint32_t vbits = 0;       // 8 validity bits per in-flight chunk
        for (int ch = -D; ch < nchunks; ++ch) {
            const int ci = ch + D;
            if (ci < nchunks) {
                const uint32_t slot = sbase + L::RAW_OFF + (ci % RD) * L::RAW_BYTES + tid * L::SLOT;
                const bool pok = p_i < P;
                const float *img = a.x + (size_t)n_i * H * W * Cin;
                const int h0 = ho_i * stride, w0 = wo_i * stride;
                uint32_t vm = 0;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const int hi = h0 + g_dr[g], wi = w0 + g_ds[g];
                    const bool v = pok && g_ok[g] && (unsigned)hi < (unsigned)H && (unsigned)wi < (unsigned)W;
                    const float *src = v ? img + ((size_t)hi * W + wi) * Cin + g_ch[g] : a.x;
                    if (use_ca) cp_async16_ca(slot + g * 16, src, v ? 16u : 0u);
                    else cp_async16(slot + g * 16, src, v ? 16u : 0u);
                    vm |= (uint32_t)v << g;
                }
#pragma unroll
                for (int g = 0; g < NPB; ++g)
                    cp_async16(slot + (8 + g) * 16, pok ? a.dy + (size_t)p_i * Cout + o0 + q * (BN / 4) + g * 4 : a.dy, pok ? 16u : 0u);
                const uint32_t sh = 8 * (ci % RD);
                vbits = (vbits & ~(0xFFu << sh)) | (vm << sh);
                p_i += 32; wo_i += 32;
                while (wo_i >= Wo) { wo_i -= Wo; if (++ho_i == Ho) { ho_i = 0; ++n_i; } }
            }
            cp_async_commit();
            if (ch < 0) continue;
            cp_async_wait<D>();
            const unsigned char *slot = smem + L::RAW_OFF + (ch % RD) * L::RAW_BYTES + tid * L::SLOT;
            const uint32_t stage = ch % NST, phase = (ch / NST) & 1;
            const uint32_t vm = (vbits >> (8 * (ch % RD))) & 0xFFu;
            if (lane == 0) mbar_wait(bar(NST + stage), phase ^ 1);
            __syncwarp();
            unsigned char *sA = smem + stage * L::STAGE_BYTES + q * 4096 + rowoff;       // column block q
            unsigned char *sB = smem + stage * L::STAGE_BYTES + L::A_BYTES + rowoff;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float4 xv = *reinterpret_cast<const float4 *>(slot + g * 16);
                if (pro) {
                    const float4 sc = *reinterpret_cast<const float4 *>(s_scale + g_ch[g]);
                    const float4 sf = *reinterpret_cast<const float4 *>(s_shift + g_ch[g]);
                    xv.x = fmaf(xv.x, sc.x, sf.x); xv.y = fmaf(xv.y, sc.y, sf.y);
                    xv.z = fmaf(xv.z, sc.z, sf.z); xv.w = fmaf(xv.w, sc.w, sf.w);
                    if (relu) { xv.x = fmaxf(xv.x, 0.f); xv.y = fmaxf(xv.y, 0.f); xv.z = fmaxf(xv.z, 0.f); xv.w = fmaxf(xv.w, 0.f); }
                }
                if (!((vm >> g) & 1u)) xv = make_float4(0.f, 0.f, 0.f, 0.f);
                const uint32_t off = (g ^ sw) << 4;
                uint4 h;
                h.x = to_tf32(xv.x); h.y = to_tf32(xv.y); h.z = to_tf32(xv.z); h.w = to_tf32(xv.w);
                *reinterpret_cast<uint4 *>(sA + off) = h;
                if (PASSES > 1) {
                    uint4 l;
                    l.x = to_tf32(xv.x - __uint_as_float(h.x)); l.y = to_tf32(xv.y - __uint_as_float(h.y));
                    l.z = to_tf32(xv.z - __uint_as_float(h.z)); l.w = to_tf32(xv.w - __uint_as_float(h.w));
                    *reinterpret_cast<uint4 *>(sA + 4 * 4096 + off) = l;
                }
            }
#pragma unroll
            for (int g = 0; g < NPB; ++g) {
                const float4 bv = *reinterpret_cast<const float4 *>(slot + (8 + g) * 16);
                dbp[g * 4] += bv.x; dbp[g * 4 + 1] += bv.y; dbp[g * 4 + 2] += bv.z; dbp[g * 4 + 3] += bv.w;
                const int pc = q * NPB + g;                 // piece index along the channel axis (4 channels each)
                const uint32_t off = (pc >> 3) * 4096 + (((pc & 7) ^ sw) << 4);
                uint4 h;
                h.x = to_tf32(bv.x); h.y = to_tf32(bv.y); h.z = to_tf32(bv.z); h.w = to_tf32(bv.w);
                *reinterpret_cast<uint4 *>(sB + off) = h;
                if (PASSES > 1) {
                    uint4 l;
                    l.x = to_tf32(bv.x - __uint_as_float(h.x)); l.y = to_tf32(bv.y - __uint_as_float(h.y));
                    l.z = to_tf32(bv.z - __uint_as_float(h.z)); l.w = to_tf32(bv.w - __uint_as_float(h.w));
                    *reinterpret_cast<uint4 *>(sB + (BNP / 32) * 4096 + off) = l;
                }
            }
            fence_proxy_async();
            mbar_arrive(bar(stage));
        }
        if (a.db != nullptr && mt == 0) {
#pragma unroll
            for (int i = 0; i < NPB * 4; ++i) {
                float t = warp_sum(dbp[i]);
                if (j == 0) atomicAdd(&a.db[o0 + q * (BN / 4) + i], t);
            }
        }
    } else if (warp == 8) {
        if (nchunks > 0) {      // converged warp, elected issue
            // c_format F32, a/b TF32, a_major = b_major = MN (bits 15, 16), N = BNP, M = 128
            constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                       ((uint32_t)(BNP >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            for (int ch = 0; ch < nchunks; ++ch) {
                const uint32_t stage = ch % NST, phase = (ch / NST) & 1;
                mbar_wait(bar(stage), phase);
                __syncwarp();
                tc_fence_after();
                const uint32_t sa = sbase + stage * L::STAGE_BYTES;
                const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {            // 4 groups of 8 pixels
                    const uint64_t ah = make_desc_mn(sa + ks * a.kstep, a.lbo16, a.sbo16), bh = make_desc_mn(sb + ks * a.kstep, a.lbo16, a.sbo16);
                    const uint32_t first = (ch == 0 && ks == 0) ? 0u : 1u;
                    if (PASSES > 1) {
                        const uint64_t al = make_desc_mn(sa + 4 * 4096 + ks * a.kstep, a.lbo16, a.sbo16);
                        const uint64_t bl = make_desc_mn(sb + (BNP / 32) * 4096 + ks * a.kstep, a.lbo16, a.sbo16);
                        mma_tf32(tmem_base, ah, bl, IDESC, first);
                        mma_tf32(tmem_base, al, bh, IDESC, 1u);
                        mma_tf32(tmem_base, ah, bh, IDESC, 1u);
                    } else {
                        mma_tf32(tmem_base, ah, bh, IDESC, first);
                    }
                }
                mma_commit(bar(NST + stage));
                if (ch == nchunks - 1) mma_commit(bar(2 * NST));
            }
        }
    } else if (nchunks > 0) {
        const int ew = warp - 4;
        const int kd = kd0 + ew * 32 + lane;
        if (lane == 0) mbar_wait(bar(2 * NST), 0);
        __syncwarp();
        tc_fence_after();
#pragma unroll
        for (int cb = 0; cb < BN; cb += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(ew * 32) << 16) + cb, v);
            if (kd < Kw) {
                float *dst = a.dw + (size_t)kd * a.Cout + o0 + cb;
#pragma unroll
                for (int t = 0; t < 16; t += 4) red_add_v4(dst + t, v[t], v[t + 1], v[t + 2], v[t + 3]);
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

template <int BN, int PASSES>
int launch(WArgs &a, cudaStream_t st) {
    using L = Lay<BN, PASSES>;
    static bool done = false;
    if (!done) {
        if (cudaFuncSetAttribute(k_wgrad_mn<BN, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess) return -1;
        done = true;
    }
    const int Kw = a.k * a.k * a.Cin;
    a.mtiles = (Kw + TM - 1) / TM;
    a.ntiles = a.Cout / BN;
    const int tiles = a.mtiles * a.ntiles;
    const int P = a.N * a.Ho * a.Wo;
    const int total_chunks = (P + 31) / 32;
    int splits = 148 / tiles; if (splits < 1) splits = 1;
    if (splits > total_chunks) splits = total_chunks;
    a.chunks_per_split = (total_chunks + splits - 1) / splits;
    a.splits = (total_chunks + a.chunks_per_split - 1) / a.chunks_per_split;
    if (launch_pdl(2, k_wgrad_mn<BN, PASSES>, dim3(tiles * a.splits), dim3(NTHREADS), L::TOTAL, st, a) != cudaSuccess) return -1;
    return 0;
}

}  // namespace

int dpp_conv2d_wgrad_tc_mn(const dpp_conv_desc *d, const float *x, const dpp_bn_ref *in_bn, const float *dy, float *dw,
                           float *db, void *stream) {
    if (d->precision != 1 && d->precision != 2) return DPP_ENOTSUP;
    if (d->Cout % 16 || d->Cout > 256 || (d->Cout > 128 && d->Cout % 128)) return DPP_ENOTSUP;
    WArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.dy = dy; a.dw = dw; a.db = db;
    a.N = d->N; a.H = d->H; a.W = d->W; a.Cin = d->Cin; a.Cout = d->Cout;
    a.k = d->k; a.stride = d->stride; a.pad = d->pad; a.Ho = d->Ho; a.Wo = d->Wo;
    if (in_bn) { a.in_bn = *in_bn; a.has_in_bn = 1; }
    { const char *e; a.lbo16 = (e = getenv("DPP_MN_LBO")) ? atoi(e) : 256; a.sbo16 = (e = getenv("DPP_MN_SBO")) ? atoi(e) : 32;
      a.kstep = (e = getenv("DPP_MN_KSTEP")) ? atoi(e) : 1024;
      a.knobs = (e = getenv("DPP_WG_KNOBS")) ? atoi(e) : 0; }
    const int bn = d->Cout > 128 ? 128 : d->Cout;
    const bool p3 = d->precision == 1;
    int rc = -1;
    if (bn == 16) rc = p3 ? launch<16, 2>(a, S(stream)) : launch<16, 1>(a, S(stream));
    else if (bn == 32) rc = p3 ? launch<32, 2>(a, S(stream)) : launch<32, 1>(a, S(stream));
    else if (bn == 64) rc = p3 ? launch<64, 2>(a, S(stream)) : launch<64, 1>(a, S(stream));
    else if (bn == 128) rc = p3 ? launch<128, 2>(a, S(stream)) : launch<128, 1>(a, S(stream));
    if (rc != 0) return dpp::fail(DPP_ECUDA, "%s: launch setup failed", __func__);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}
