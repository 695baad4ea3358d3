// ABI plumbing: error state, device info, layout conversion, fills.
#include "common.cuh"

#include <stdlib.h>

namespace dpp {
thread_local char g_err[512] = "";
static int g_pdl = -1;
int pdl_mode() {
    if (g_pdl < 0) { const char *e = getenv("DPP_PDL"); g_pdl = e ? atoi(e) : DPP_PDL_DEFAULT; }
    return g_pdl;
}
}

using namespace dpp;

extern "C" int dpp_abi_version(void) { return DPP_ABI_VERSION; }
extern "C" const char *dpp_last_error(void) { return dpp::g_err; }
extern "C" int dpp_set_pdl(int mode) { dpp::g_pdl = mode < 0 ? -1 : mode; return DPP_OK; }

extern "C" int dpp_device_info(int device, int *cc_out, int *sm_count_out, char *name_out, int name_cap) {
    cudaDeviceProp p;
    DPP_CUDA(cudaGetDeviceProperties(&p, device));
    if (cc_out) *cc_out = p.major * 10 + p.minor;
    if (sm_count_out) *sm_count_out = p.multiProcessorCount;
    if (name_out && name_cap > 0) {
        strncpy(name_out, p.name, name_cap - 1);
        name_out[name_cap - 1] = 0;
    }
    return DPP_OK;
}

// NCHW -> NHWC through a 32x32 shared-memory transpose of the (C, H*W) plane.
__global__ void k_transpose_planes(const float *__restrict__ src, float *__restrict__ dst, int rows, int cols) {
    // per image: src [rows][cols] -> dst [cols][rows]
    __shared__ float tile[32][33];
    const float *s = src + (size_t)blockIdx.z * rows * cols;
    float *d = dst + (size_t)blockIdx.z * rows * cols;
    int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = s[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) d[(size_t)c * rows + r] = tile[threadIdx.x][i];
    }
}

static int transpose_planes(const float *src, float *dst, int N, int rows, int cols, void *stream) {
    dim3 grid(cdiv(cols, 32), cdiv(rows, 32), N), block(32, 8);
    k_transpose_planes<<<grid, block, 0, S(stream)>>>(src, dst, rows, cols);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

extern "C" int dpp_nchw_to_nhwc(const float *src, float *dst, int N, int C, int H, int W, void *stream) {
    DPP_CHECK_ARG(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && N <= 65535);
    if (C == 1) {
        DPP_CUDA(cudaMemcpyAsync(dst, src, sizeof(float) * (size_t)N * H * W, cudaMemcpyDeviceToDevice, S(stream)));
        return DPP_OK;
    }
    return transpose_planes(src, dst, N, C, H * W, stream);
}

extern "C" int dpp_nhwc_to_nchw(const float *src, float *dst, int N, int C, int H, int W, void *stream) {
    DPP_CHECK_ARG(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && N <= 65535);
    if (C == 1) {
        DPP_CUDA(cudaMemcpyAsync(dst, src, sizeof(float) * (size_t)N * H * W, cudaMemcpyDeviceToDevice, S(stream)));
        return DPP_OK;
    }
    return transpose_planes(src, dst, N, H * W, C, stream);
}

template <typename T>
__global__ void k_fill(T *p, T v, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

extern "C" int dpp_fill_f32(float *p, float value, int64_t n, void *stream) {
    DPP_CHECK_ARG(p && n >= 0);
    if (n == 0) return DPP_OK;
    if (value == 0.f) {
        DPP_CUDA(cudaMemsetAsync(p, 0, sizeof(float) * n, S(stream)));
        return DPP_OK;
    }
    int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    k_fill<float><<<blocks, 256, 0, S(stream)>>>(p, value, n);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

extern "C" int dpp_fill_f64(double *p, double value, int64_t n, void *stream) {
    DPP_CHECK_ARG(p && n >= 0);
    if (n == 0) return DPP_OK;
    if (value == 0.0) {
        DPP_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * n, S(stream)));
        return DPP_OK;
    }
    int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    k_fill<double><<<blocks, 256, 0, S(stream)>>>(p, value, n);
    DPP_LAUNCH_CHECK();
    return DPP_OK;
}

// Strided device-to-device copy (rows x width bytes): concatenation of flattened tower outputs into the first
// hidden layer's input (reference: net/scalenet.py:174-178, T.concatenate(..., axis=1)).
extern "C" int dpp_copy2d(void *dst, int64_t dst_pitch, const void *src, int64_t src_pitch, int64_t width, int64_t rows,
                          void *stream) {
    DPP_CHECK_ARG(dst && src && width > 0 && rows > 0 && dst_pitch >= width && src_pitch >= width);
    DPP_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width, (size_t)rows,
                               cudaMemcpyDeviceToDevice, S(stream)));
    return DPP_OK;
}
