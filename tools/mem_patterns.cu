// Micro-benchmark (B200): sustained global-store and TMA-load throughput of the access patterns the HiddenLayer GEMMs
// produce - few warps per SM, 128-byte or 512-byte pieces, 4 KB row pitch.  Decides the epilogue / producer layout of
// csrc/fc_stream.cu.  Build + run:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mem_patterns
// tools/mem_patterns.cu && tools/mem_patterns
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

// ---- stores: matrix [rows][1024] floats (4 KB pitch).  A "tile" = 128 rows x 128 columns (512 B per row).
// mode 0: warp w of 4 writes columns [32w, 32w+32) of one row per instruction (128 B), rows in sequence (the dW epilogue)
// mode 1: warp writes a whole 512-byte tile row per instruction (st.v4), 4 warps -> 4 rows per round
// mode 2: like 0 but the CTA's tile is stored as a contiguous 64 KB block (workspace pattern)
__global__ void __launch_bounds__(512) k_store(float *C, int rows, int mode, int warps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= warps) return;
    const int tiles = (rows / 128) * 8;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int mt = t & 7, nt = t >> 3;
        if (mode == 0) {
            for (int w4 = warp; w4 < 4 * (warps > 4 ? 1 : 1) && w4 < 4; w4 += warps) {
                float *dst = C + (size_t)(nt * 128) * 1024 + mt * 128 + w4 * 32 + lane;
#pragma unroll 16
                for (int r = 0; r < 128; ++r) dst[(size_t)r * 1024] = (float)r;
            }
        } else if (mode == 1) {
            float4 *dst = reinterpret_cast<float4 *>(C + (size_t)(nt * 128) * 1024 + mt * 128) + lane;
#pragma unroll 8
            for (int r = warp; r < 128; r += warps) dst[(size_t)r * 256] = make_float4(1.f, 2.f, 3.f, (float)r);
        } else if (mode == 2) {
            for (int w4 = warp; w4 < 4; w4 += warps) {
                float *dst = C + (size_t)t * 16384 + w4 * 32 + lane;
#pragma unroll 16
                for (int r = 0; r < 128; ++r) dst[r * 128] = (float)r;
            }
        } else {   // mode 3: 16 warps, warp = (row group, column quarter): 4 rows in flight per quarter
            const int w4 = warp & 3, rg = warp >> 2, nrg = warps >> 2;
            float *dst = C + (size_t)(nt * 128) * 1024 + mt * 128 + w4 * 32 + lane;
#pragma unroll 8
            for (int r = rg; r < 128; r += nrg) dst[(size_t)r * 1024] = (float)r;
        }
    }
}

// ---- TMA loads: one thread per CTA streams boxes of a [rows][1024] float matrix into a 4-slot ring and waits.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128) k_tma(const __grid_constant__ CUtensorMap tm, int rows, int mode, int depth, int *sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[8];
    const uint32_t sb = (smem_u32(smem) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    // mode 0: box {32 k, 128 rows} (128-byte pieces, the dx pattern): CTA owns 128 rows, walks k
    // mode 1: box {128 cols, 32 rows} (512-byte pieces, the forward pattern): CTA owns a 128-column strip x row range
    const uint64_t tmap = reinterpret_cast<uint64_t>(&tm);
    int issued = 0, waited = 0, total;
    int c0s, c1s;
    if (mode == 0) total = (rows / 128 + gridDim.x - 1 - blockIdx.x) / gridDim.x * 32;      // row tiles of this CTA x 32 k-chunks
    else total = rows / 32 / (gridDim.x / 8);                                               // chunks of 32 rows
    for (; waited < total;) {
        while (issued < total && issued - waited < depth) {
            const int slot = issued % depth;
            const uint32_t bar = smem_u32(&bars[slot]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16384;" ::"r"(bar) : "memory");
            if (mode == 0) { const int tile = blockIdx.x + (issued / 32) * gridDim.x; c0s = (issued % 32) * 32; c1s = tile * 128; }
            else { const int strip = blockIdx.x & 7, part = blockIdx.x >> 3; c0s = strip * 128; c1s = (part * total + issued) * 32; }
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                             sb + slot * 16384),
                         "l"(tmap), "r"(c0s), "r"(c1s), "r"(bar)
                         : "memory");
            ++issued;
        }
        const int slot = waited % depth;
        const uint32_t bar = smem_u32(&bars[slot]), parity = (waited / depth) & 1;
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok)
                         : "r"(bar), "r"(parity)
                         : "memory");
        } while (!ok);
        ++waited;
    }
    if (sink != nullptr && rows < 0) *sink = issued;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int rows = 65536;                      // 65536 x 1024 floats = 268 MB: larger than the 126 MB L2
    float *C;
    cudaMalloc(&C, (size_t)rows * 1024 * 4);
    cudaMemset(C, 0, (size_t)rows * 1024 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *sn[4] = {"128 B per warp-store, 4 KB pitch (dW epilogue)", "512 B per warp-store (st.v4), 4 KB pitch",
                         "128 B per warp-store, contiguous 64 KB tiles", "128 B per warp-store, 4 KB pitch, rows spread over warps"};
    for (int mode = 0; mode < 4; ++mode)
        for (int warps = 4; warps <= 16; warps *= 2) {
            if (mode == 3 && warps == 4) continue;
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                k_store<<<148, 512>>>(C, rows, mode, warps);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("store mode %d (%s), %2d warps/SM: %.1f us -> %.0f GB/s\n", mode, sn[mode], warps, best * 1e3, 268.4 / best);
        }
    // TMA
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);
    cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 1024);
    for (int mode = 0; mode < 2; ++mode)
        for (int promo = 0; promo < 2; ++promo)
            for (int depth = 2; depth <= 8; depth *= 2) {
                CUtensorMap tm;
                cuuint64_t gdim[2] = {1024, (cuuint64_t)rows}, gstr[1] = {4096};
                cuuint32_t box[2], estr[2] = {1, 1};
                if (mode == 0) { box[0] = 32; box[1] = 128; } else { box[0] = 128; box[1] = 32; }
                CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, C, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 mode == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                 promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
                const int grid = mode == 0 ? 128 : 144;
                float best = 1e9f;
                for (int rep = 0; rep < 3; ++rep) {
                    cudaEventRecord(e0);
                    k_tma<<<grid, 128, 8 * 16384 + 1024>>>(tm, rows, mode, depth, nullptr);
                    cudaEventRecord(e1);
                    cudaError_t e = cudaEventSynchronize(e1);
                    if (e != cudaSuccess) { printf("k_tma: %s\n", cudaGetErrorString(e)); return 1; }
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    if (ms < best) best = ms;
                }
                const double mb = mode == 0 ? 268.4 : 268.4 * (double)(rows / 32 / 18 * 18) / (rows / 32);
                printf("TMA mode %d (%s), L2 promotion %s, %d boxes (16 KB) in flight per SM, %d CTAs: %.1f us -> %.0f GB/s\n", mode,
                       mode == 0 ? "box 32 k x 128 rows: 128-byte pieces at 4 KB pitch" : "box 128 cols x 32 rows: 512-byte pieces at 4 KB pitch",
                       promo ? "256B" : "128B", depth, grid, best * 1e3, mb / best);
            }
    return 0;
}
