// tcgen05 implicit-GEMM convolution (precision modes 1 = 3xTF32, 2 = TF32).  Placeholder until the
// tensor-core path lands: reports "not supported" so dpp_conv2d_fwd falls through to fp32 SIMT.
#include "common.cuh"
int dpp_conv2d_fwd_tc(const dpp_conv_desc *, const float *, const dpp_bn_ref *, const float *, const float *,
                      const float *, float *, double *, void *) {
    return DPP_ENOTSUP;
}
