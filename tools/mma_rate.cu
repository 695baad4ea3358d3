// Micro-benchmark: issue rate of tcgen05.mma on one SM (B200), cycles per instruction for back-to-back MMAs
// into one accumulator.  kind::tf32 (K = 8) and kind::f16 with bf16 operands (K = 16), A from shared memory (.ss)
// or tensor memory (.ts), M = 128, N = 16 .. 256.  Build + run:  make -C tools mma_rate && tools/mma_rate
// Used to decide whether the 3xTF32 kernels (conv, backward-weights, FC) are bound by the tensor pipe.
#include "../deep-prior-pp_b200/csrc/tc_common.cuh"
#include <cstdio>
#include <cstdlib>

namespace dpp { thread_local char g_err[512]; }
using namespace dpp::tc;

__device__ __forceinline__ void mma_f16_ss_1t(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

// mode 0: tf32 .ss, 1: tf32 .ts, 2: bf16 .ss
__global__ void __launch_bounds__(128, 1) k_rate(int mode, int N, int reps, long long *out) {
    extern __shared__ unsigned char raw[];
    unsigned char *smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    __shared__ uint32_t slot;
    __shared__ __align__(8) unsigned long long barmem;
    const uint32_t bar = smem_u32(&barmem);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x < 32) {
        if (elect_one()) {
            const uint64_t a0 = make_desc(sbase), b0 = make_desc(sbase + 16384);
            // idesc: D fp32 (1<<4); A/B format at bits 7 / 10: tf32 = 2, bf16 = 1; N>>3 at 17, M>>4 at 24
            const uint32_t fmt = mode == 2 ? 1u : 2u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int rep = 0; rep < 3; ++rep) {
                const long long t0 = clock64();
                for (int i = 0; i < reps; ++i) {
                    const uint64_t off = 2 * (i & 3);
                    if (mode == 0) mma_tf32_ss_1t(tm, a0 + off, b0 + off, idesc, 1u);
                    else if (mode == 1) mma_tf32_ts_1t(tm, tm + 256 + 8 * (i & 3), b0 + off, idesc, 1u);
                    else mma_f16_ss_1t(tm, a0 + off, b0 + off, idesc, 1u);
                }
                mma_commit_1t(bar);
                const long long t1 = clock64();
                mbar_wait(bar, rep & 1);
                const long long t2 = clock64();
                out[0] = t1 - t0; out[1] = t2 - t0;
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
    }
}

int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    const int SMEM = 16384 + 32768 + 1024;
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    const char *names[3] = {"tf32.ss", "tf32.ts", "bf16.ss"};
    const int reps = 512;
    for (int mode = 0; mode < 3; ++mode)
        for (int N = 16; N <= 256; N *= 2) {
            k_rate<<<1, 128, SMEM>>>(mode, N, reps, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s N=%d: %s\n", names[mode], N, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            const double k = mode == 2 ? 16 : 8;
            printf("%s M=128 N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA -> %.0f flop/clk/SM\n", names[mode], N,
                   (double)h[0] / reps, (double)h[1] / reps, 2.0 * 128 * N * k * reps / (double)h[1]);
        }
    return 0;
}
