// Host execution of the per-thread core of csrc/convpool8.cu (convpool8.cuh is shared between the device kernel and
// this file): emulates the kernel's tile loop on the CPU and compares outputs and arg-max cells, bit for bit, with
// an independent direct convolution + max-pool.  TEST INFRASTRUCTURE ONLY; built and run by tests/test_host_convpool8.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../deep-prior-pp_b200/csrc/convpool8.cuh"

using namespace dpp;

static unsigned long long rng_state = 88172645463325252ull;
static float frand() {                       // xorshift64*, uniform in [-1, 1)
    rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
    return (float)((rng_state * 2685821657736338717ull >> 40) / 8388608.0 - 1.0);
}

// independent reference: full-resolution 'valid' convolution (taps r, s, channels c, fused multiply-add), then
// non-overlapping max-pool (first maximum wins), bias, ReLU.  x NHWC, w [(r*K+s)*CIN+c][8], y [N][Hp][Wp][8].
static void reference(const std::vector<float> &x, const std::vector<float> &w, const float *bias, int N, int H, int W,
                      int K, int CIN, int POOL, int relu, std::vector<float> &y, std::vector<uint8_t> &am) {
    const int Hc = H - K + 1, Wc = W - K + 1, Hp = Hc / POOL, Wp = Wc / POOL;
    std::vector<float> conv((size_t)Hc * Wc * 8);
    y.assign((size_t)N * Hp * Wp * 8, 0.f);
    am.assign((size_t)N * Hp * Wp * 8, 0);
    for (int n = 0; n < N; ++n) {
        for (int oy = 0; oy < Hc; ++oy)
            for (int ox = 0; ox < Wc; ++ox)
                for (int q = 0; q < 8; ++q) {
                    float a = 0.f;
                    for (int r = 0; r < K; ++r)
                        for (int s = 0; s < K; ++s)
                            for (int c = 0; c < CIN; ++c)
                                a = fmaf(x[(((size_t)n * H + oy + r) * W + ox + s) * CIN + c],
                                         w[((size_t)(r * K + s) * CIN + c) * 8 + q], a);
                    conv[((size_t)oy * Wc + ox) * 8 + q] = a;
                }
        for (int ph = 0; ph < Hp; ++ph)
            for (int pw = 0; pw < Wp; ++pw)
                for (int q = 0; q < 8; ++q) {
                    float best = -INFINITY;
                    uint8_t cell = 0;
                    for (int cy = 0; cy < POOL; ++cy)
                        for (int cx = 0; cx < POOL; ++cx) {
                            float v = conv[((size_t)(ph * POOL + cy) * Wc + pw * POOL + cx) * 8 + q];
                            if (v > best) { best = v; cell = (uint8_t)(cy * POOL + cx); }
                        }
                    float v = best + bias[q];
                    if (relu) v = fmaxf(v, 0.f);
                    size_t o = (((size_t)n * Hp + ph) * Wp + pw) * 8 + q;
                    y[o] = v;
                    am[o] = cell;
                }
    }
}

// the kernel's tile loop (k_convpool8_fwd), one "thread" after the other
template <int K, int CIN, int POOL>
static int run_case(int N, int H, int W, int relu) {
    constexpr int TP8 = 16, P = TP8 * POOL + K - 1;
    const int Hp = (H - K + 1) / POOL, Wp = (W - K + 1) / POOL;
    const int tilesY = (Hp + TP8 - 1) / TP8, tilesX = (Wp + TP8 - 1) / TP8;
    std::vector<float> x((size_t)N * H * W * CIN), w((size_t)K * K * CIN * 8), patch((size_t)P * P * CIN);
    float bias[8];
    for (auto &v : x) v = frand();
    for (auto &v : w) v = frand() * 0.3f;
    for (auto &v : bias) v = frand() * 0.1f;
    for (size_t i = 0; i < x.size(); i += 97) x[i] = x[(i + 1) % x.size()];      // some exact ties for the arg-max rule
    std::vector<float> y((size_t)N * Hp * Wp * 8, -7.f), yr;
    std::vector<uint8_t> am((size_t)N * Hp * Wp * 8, 255), amr;
    for (int tile = 0; tile < N * tilesY * tilesX; ++tile) {
        const int n = tile / (tilesY * tilesX), tr = tile % (tilesY * tilesX);
        const int ty0 = tr / tilesX, tx0 = tr % tilesX;
        const int y0 = ty0 * TP8 * POOL, x0 = tx0 * TP8 * POOL;
        for (int i = 0; i < P * P * CIN; ++i) {
            const int c = i % CIN, pp = i / CIN, px = pp % P, py = pp / P, yy = y0 + py, xx = x0 + px;
            patch[i] = (yy < H && xx < W) ? x[(((size_t)n * H + yy) * W + xx) * CIN + c] : 0.f;
        }
        for (int t = 0; t < TP8 * TP8; ++t) {
            const int ly = t / TP8, lx = t % TP8, ph = ty0 * TP8 + ly, pw = tx0 * TP8 + lx;
            if (ph < Hp && pw < Wp) {
                float best[8];
                uint8_t bidx[8];
                convpool8_pixel<K, CIN, POOL>(patch.data(), P, w.data(), ly, lx, best, bidx);
                convpool8_store(best, bidx, bias, relu, y.data(), am.data(), (((size_t)n * Hp + ph) * Wp + pw) * 8);
            }
        }
    }
    reference(x, w, bias, N, H, W, K, CIN, POOL, relu, yr, amr);
    size_t bad = 0;
    for (size_t i = 0; i < y.size(); ++i)
        if (memcmp(&y[i], &yr[i], 4) != 0 || am[i] != amr[i]) ++bad;
    printf("K=%d CIN=%d POOL=%d H=%d -> %dx%d: %zu of %zu values differ\n", K, CIN, POOL, H, Hp, Wp, bad, y.size());
    return bad == 0 ? 0 : 1;
}

int main() {
    int rc = 0;
    rc |= run_case<5, 1, 4>(2, 128, 128, 1);     // ScaleNet tower 1 / PoseRegNet layer 1
    rc |= run_case<5, 8, 2>(2, 31, 31, 1);
    rc |= run_case<3, 8, 1>(2, 13, 13, 1);
    rc |= run_case<5, 1, 2>(2, 64, 64, 1);       // tower 2
    rc |= run_case<5, 8, 2>(2, 30, 30, 1);
    rc |= run_case<5, 1, 2>(2, 32, 32, 1);       // tower 3
    rc |= run_case<5, 8, 1>(2, 14, 14, 1);
    rc |= run_case<3, 8, 1>(3, 10, 10, 0);
    rc |= run_case<5, 1, 4>(1, 131, 77, 0);      // ragged tiles, non-square
    printf(rc == 0 ? "convpool8 host test OK\n" : "convpool8 host test FAILED\n");
    return rc;
}
