// Per-thread core of the 8-filter conv + max-pool kernels (ScaleNet / PoseRegNet towers), shared between the device
// kernel (convpool8.cu) and a host test (tests/convpool8_host_test.cu) that runs the very same code on the CPU.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CP8_HD __host__ __device__ __forceinline__
#else
#define CP8_HD inline
#endif

namespace dpp {

// One pooled output pixel (ly, lx) of a tile, all 8 output channels.
//   patch: the tile's input patch [P][P][CIN] (P = TP * POOL + K - 1), pixel-major, channels contiguous
//   ws   : weights [(r * K + s) * CIN + c][8]   (the library's "KC" layout, already flipped)
// Accumulation order per output value: taps r, s, then input channels c, fused multiply-adds - the order of the
// generic kernel (convpool.cu::k_convpool_fwd), so results are bit-identical to it.  Pool cells are visited row by
// row; the first maximum wins (bidx = cy * POOL + cx), like the generic kernel and the backward pass expect.
template <int K, int CIN, int POOL>
CP8_HD void convpool8_pixel(const float *patch, int P, const float *ws, int ly, int lx, float best[8], uint8_t bidx[8]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) { best[q] = -INFINITY; bidx[q] = 0; }
    for (int cy = 0; cy < POOL; ++cy) {
        float acc[POOL][8];
#pragma unroll
        for (int cx = 0; cx < POOL; ++cx)
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[cx][q] = 0.f;
        const float *row0 = patch + ((ly * POOL + cy) * P + lx * POOL) * CIN;
        // unrolling: all input channels of a tap (and, for CIN == 1, all taps of a filter row); unrolling the whole
        // K*K*CIN nest makes ptxas hoist hundreds of shared-memory loads (255 registers, kilobytes of spills)
#pragma unroll 1
        for (int r = 0; r < K; ++r) {
#pragma unroll(CIN == 1 ? K : 1)
            for (int s = 0; s < K; ++s) {
#pragma unroll
                for (int c = 0; c < CIN; ++c) {
                    const float *wp = ws + ((r * K + s) * CIN + c) * 8;
                    float w8[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) w8[q] = wp[q];
#pragma unroll
                    for (int cx = 0; cx < POOL; ++cx) {
                        const float xv = row0[(r * P + s + cx) * CIN + c];
#pragma unroll
                        for (int q = 0; q < 8; ++q) acc[cx][q] = fmaf(xv, w8[q], acc[cx][q]);
                    }
                }
            }
        }
#pragma unroll
        for (int cx = 0; cx < POOL; ++cx) {
            const uint8_t cell = (uint8_t)(cy * POOL + cx);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (acc[cx][q] > best[q]) { best[q] = acc[cx][q]; bidx[q] = cell; }
        }
    }
}

// bias, activation and the store of one pooled pixel (shared so that host test and kernel agree on the layout)
CP8_HD void convpool8_store(const float best[8], const uint8_t bidx[8], const float *bias, int relu, float *y,
                            uint8_t *argmax, size_t ob) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float v = best[q] + bias[q];
        if (relu) v = fmaxf(v, 0.f);
        y[ob + q] = v;
        if (argmax) argmax[ob + q] = bidx[q];
    }
}

}  // namespace dpp
