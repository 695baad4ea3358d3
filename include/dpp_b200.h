/*
 * dpp_b200.h - C ABI of the B200-native DeepPrior++ hot path (libdpp_b200.so).
 *
 * The reference (moberweger/deep-prior-pp) has no C/FFI boundary for this path: its
 * boundary is the Python class surface (net.NetBase / trainer.PoseRegNetTrainer) that
 * compiles Theano graphs.  This header is the boundary the replacement puts UNDER that
 * Python surface: every Theano op call site on the hot path becomes one entry point here
 * (loaded with ctypes; see INTEGRATION.md for the reference-side stubs).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  All pointers are DEVICE pointers unless
 *     the name ends in _host.  The caller owns every buffer; the library allocates nothing.
 *   - `stream` is a cudaStream_t passed as void*.  Calls are asynchronous on that stream
 *     and capturable into a CUDA graph (no allocation, no synchronisation inside).
 *   - Return value: 0 = OK, negative = DPP_E*.  dpp_last_error() gives the message of the
 *     last failure on the calling thread.  Nothing throws or aborts.
 *   - Activations are NHWC fp32 inside the library ("pixel-major": a pixel's channels are
 *     contiguous).  dpp_nchw_to_nhwc / dpp_nhwc_to_nchw convert at the reference boundary.
 *   - Convolution weights are stored "KC": W[(r*kw+s)*Cin + c][o], already FLIPPED, i.e.
 *     W_kc[(r,s,c)][o] = W_theano[o][c][kh-1-r][kw-1-s] (theano conv2d is a true
 *     convolution, reference net/convlayer.py:230-235).  FC weights are (n_in, n_out)
 *     row-major exactly like the reference (net/hiddenlayer.py:124,136).
 *   - BatchNorm statistics travel as fp64 sums {sum[C], sumsq[C]} produced by the
 *     epilogue of the kernel that writes the tensor (net/batchnormlayer.py:154-155 is
 *     fused into its producer) and are consumed by the prologue of the next kernel
 *     (net/batchnormlayer.py:192 + net/nonlinearitylayer.py:119 fused into the consumer).
 */
#ifndef DPP_B200_H
#define DPP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPP_OK 0
#define DPP_EINVAL (-1)   /* bad argument / unsupported shape */
#define DPP_ECUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define DPP_ENOTSUP (-3)  /* configuration not implemented by this build */

#define DPP_ABI_VERSION 2

int dpp_abi_version(void);
const char *dpp_last_error(void);
/* name of the device the library would run on, compute capability major*10+minor; fails
 * with DPP_ECUDA when no CUDA device is present (the product has no CPU path). */
int dpp_device_info(int device, int *cc_out, int *sm_count_out, char *name_out, int name_cap);
/* Programmatic dependent launch of the step's kernel chain (bit 0: main chain, bit 1: backward-weights
 * kernels); -1 = re-read DPP_PDL from the environment.  Takes effect for launches issued after the call. */
int dpp_set_pdl(int mode);
/* Allocates the library-owned workspace of the backward-weights split reduction (16 MB, idempotent).  Call once
 * outside CUDA-graph capture; dpp_conv2d_wgrad allocates it lazily otherwise. */
int dpp_wgrad_workspace_init(void);

/* ---------------------------------------------------------------------------------------
 * BatchNorm reference handed to conv/fc prologues (reference: net/batchnormlayer.py:154-192).
 * If sums != NULL: train mode, batch statistics: mean = sum/count,
 *   var = sumsq/count - mean^2 (fp64), inv_std = 1/sqrt(var + eps).
 * else: deterministic mode, the stored running mean / inv_std are used.
 * A prologue computes  a = max((x - mean) * (gamma*inv_std) + beta, 0)  when relu != 0.
 * ------------------------------------------------------------------------------------- */
typedef struct dpp_bn_ref {
    const double *sums;    /* [2*C]: sum then sumsq, or NULL */
    const float *mean;     /* [C] running mean (used when sums == NULL) */
    const float *inv_std;  /* [C] running inv_std (used when sums == NULL) */
    const float *gamma;    /* [C] */
    const float *beta;     /* [C] */
    double count;          /* N*H*W of the normalised tensor */
    float eps;             /* 1e-4 in the reference */
    int relu;              /* apply max(.,0) after the affine */
} dpp_bn_ref;

/* ---- layout conversion at the reference boundary (NCHW numpy arrays) ---------------- */
int dpp_nchw_to_nhwc(const float *src, float *dst, int N, int C, int H, int W, void *stream);
int dpp_nhwc_to_nchw(const float *src, float *dst, int N, int C, int H, int W, void *stream);

/* ---- depth-crop augmentation ---------------------------------------------------------
 * Replaces NetTrainer.augmentCrop's pixel work (trainer/nettrainer.py:948-995) and the
 * cv2 warps called from HandDetector.rotateHand / recropHand
 * (util/handdetector.py:730-738, :791-801).  One record per output sample; the host side
 * (util/handdetector.py mirror) prepares it in fp64 exactly as the reference computes the
 * matrices.  mode: 0 = none, 1 = affine NN (rotateHand), 2 = perspective NN + z-threshold
 * (moveCoM / scaleHand via recropHand); +16 = raw (skip the final CoM re-normalisation:
 * the warp alone, for direct HandDetector.rotateHand / recropHand calls).
 * Bit-exact against the oracle's index rules (oracle/augment.py).                        */
typedef struct dpp_aug_rec {
    int32_t src_index;   /* row of `crops` to read (the ORIGINAL normalised crop)        */
    int32_t mode;
    float half_old;      /* f32(cube_z/2) of the stored crop     (nettrainer.py:951)     */
    float comz_old;      /* f32 com_z of the stored crop                                   */
    float zstart, zend;  /* f32 z-thresholds of recropHand (handdetector.py:795-801)      */
    float bg;            /* f32(com_z' + cube_z'/2)               (nettrainer.py:990)     */
    float lo;            /* f32(com_z' - cube_z'/2)                                         */
    float comz_new;      /* f32 com_z'                                                       */
    float half_new;      /* f32(cube_z'/2)                                                   */
    double m[9];         /* mode 1: {i00,i01,b0,i10,i11,b1,-,-,-} inverse affine (fp64)
                            mode 2: inverse homography (cv::invert of Mnew*inv(M))         */
} dpp_aug_rec;

int dpp_augment_fwd(const float *crops, const dpp_aug_rec *recs, float *out, int n_out,
                    int H, int W, void *stream);

/* ---- hand crop of the inference cascade -------------------------------------------------
 * Replaces, per frame, HandDetector.getCrop (util/handdetector.py:260-296: window + zero padding +
 * z-threshold), resizeCrop = cv2.resize INTER_NEAREST (:336-351), cropArea3D's centred paste on a
 * getNDValue() canvas (:467-476), the normalisation of refineCoM (:640-647) or of
 * RealtimeHandposePipeline.detect (util/realtimehandposepipeline.py:327-332), refineCoM's 1/2 and 1/4
 * centre crops (:656-667) and estimatePose's mirroring (:346-349).  One record per output crop; the
 * window geometry (comToBounds, :204-226) is computed on the host in fp64 as the reference does.
 * flags: bit 0 = normalise (v == 0 -> hi; (v - comz) / half), bit 1 = also clamp to [lo, hi]
 * (refineCoM; the pipeline's own clip() result is discarded by the reference), bit 2 = mirror x.
 * Bit-exact against oracle/cascade.py (cv2 4.13.0 resizeNN index rule).                          */
typedef struct dpp_crop_rec {
    int32_t src_index;      /* frame of `frames` to read                                          */
    int32_t xstart, ystart; /* top-left of the window in the frame (may be negative)              */
    int32_t wb, hb;         /* window size (xend - xstart, yend - ystart)                         */
    int32_t rw, rh;         /* cv2.resize target size                                             */
    int32_t px, py;         /* where the resized patch is pasted in the H x W output              */
    int32_t flags;
    float zstart, zend;     /* f32 z-thresholds of getCrop                                        */
    float fill;             /* canvas value outside the patch (getNDValue)                        */
    float hi, lo;           /* f32(com_z + cube_z/2), f32(com_z - cube_z/2)                       */
    float comz, half;       /* f32 com_z, f32(cube_z/2)                                           */
    float reserved;
    double ifx, ify;        /* 1. / (rw / wb), 1. / (rh / hb) in fp64 (cv2's inverse scale)       */
} dpp_crop_rec;

/* frames [n_frames, Hf, Wf] fp32 (mm); out0 [n_out, H, W]; out1 [n_out, H/2, W/2] and
 * out2 [n_out, H/4, W/4] (centre crops of out0, may be NULL).  W % 32 == 0, H % 8 == 0.          */
int dpp_recrop_fwd(const float *frames, const dpp_crop_rec *recs, float *out0, float *out1, float *out2,
                   int n_out, int Hf, int Wf, int H, int W, void *stream);

/* ---- random pose sampling for the PCA prior --------------------------------------------
 * HandDetector.sampleRandomPoses (util/handdetector.py:805-909, rot3D = False) for n poses: the host
 * draws the random parameters with the reference's NumPy stream in the reference's order (:837-841)
 * and passes them in.  mode[i]: 0 none, 1 rot, 2 sc, 3 com, 4 rot+com (= com+rot), 5 rot+com+sc
 * (= rot+sc+com); ridx[i]: row of the base arrays; off [n,3] mm (fp64); sc [n] (fp64);
 * cos_sin [n,2] = cos / sin of rot[i,0] * pi / 180 (fp64, computed by the host's libm so the result
 * is bit-identical to NumPy's).  Camera = the importer's pin-hole model (data/importers.py:80-119,
 * :756-793, :1187-1224; flip_y for NYU / MSRA15).  Outputs: new_poses [n,J,3] normalised by
 * new_cube_z / 2, new_com [n,3], new_cube [n,3].                                                  */
int dpp_sample_poses(const float *base_poses, const float *base_com, const float *base_cube,
                     const int32_t *mode, const int32_t *ridx, const double *off, const double *sc,
                     const double *cos_sin, double fx, double fy, double ux, double uy, int flip_y,
                     float *new_poses, float *new_com, float *new_cube, int n, int J, void *stream);

/* ---- pose error metrics (util/handpose_evaluation.py:92-181; trainer/poseregnettrainer.py:123-125)
 * pred, gt [n_frames, J, 3] (mm).  err [n_frames, J] = sqrt(sum((gt - pred)^2)) per joint,
 * frame_mean / frame_max [n_frames] = nan-mean / nan-max over the joints of a frame; any output may
 * be NULL.  getMeanError = mean(frame_mean), getMaxError = max(frame_max).                      */
int dpp_joint_errors(const float *pred, const float *gt, float *err, float *frame_mean, float *frame_max,
                     int n_frames, int J, void *stream);

/* ---- ConvPoolLayer (net/convpoollayer.py:251-282): conv -> maxpool -> +bias -> act ----
 * x [N,H,W,Cin] NHWC, w KC [(k*k*Cin)][Cout], y [N,Hp,Wp,Cout]; pad = k/2 ('half') or 0
 * ('valid'); pool >= 1 (floor, ignore_border).  argmax (uint8, same shape as y, may be
 * NULL for inference) records the winning pool cell for the backward pass.
 * stats (fp64 [2*Cout], may be NULL) accumulates sum/sumsq of y for a following BN.     */
int dpp_convpool_fwd(const float *x, const float *w, const float *bias, float *y,
                     uint8_t *argmax, double *stats, int N, int H, int W, int Cin, int Cout,
                     int k, int pad, int pool, int relu, void *stream);
/* dy is the gradient w.r.t. y; if relu != 0 it is masked with y > 0 first (y required).
 * dw [(k*k*Cin)][Cout] and db [Cout] are ACCUMULATED (+=); dx may be NULL (first layer). */
int dpp_convpool_bwd(const float *x, const float *w, const float *y, const uint8_t *argmax,
                     const float *dy, float *dw, float *db, float *dx, int N, int H, int W,
                     int Cin, int Cout, int k, int pad, int pool, int relu, void *stream);

/* ---- ConvLayer (net/convlayer.py:230-251) with the surrounding BN/ReLU fused ----------
 * Implicit GEMM: M = N*Ho*Wo pixels, N = Cout, K = k*k*Cin.
 * in_bn (may be NULL): BN(+ReLU) applied to x on the fly (the layers in front of the conv).
 * residual (may be NULL): added to the output (res_block's identity/shortcut sum,
 *   net/resnet.py:379,414).  out_stats (may be NULL): fp64 sum/sumsq of the written
 *   output for the next BN.
 * precision: 0 = fp32 SIMT (exact reference arithmetic order up to summation order),
 *            1 = 3xTF32 tcgen05 (fp32-equivalent), 2 = TF32 tcgen05.                    */
typedef struct dpp_conv_desc {
    int N, H, W, Cin;      /* input tensor */
    int Cout, k, stride, pad;
    int Ho, Wo;
    int precision;
    /* tcgen05 path only (precision 1|2): the layer's weights pre-packed by dpp_conv_pack_all
     * into shared-memory images; NULL selects the fp32 SIMT kernels. */
    const float *wpack_fwd;
    const float *wpack_dgrad;
} dpp_conv_desc;

/* ---- weight packing for the tcgen05 path ------------------------------------------------
 * Sizes (in floats) of the forward / dgrad images of one layer; precision as in dpp_conv_desc. */
int dpp_conv_pack_size(int Cin, int Cout, int k, int precision, int64_t *fwd_floats,
                       int64_t *dgrad_floats);
/* One launch re-packs every conv layer after the optimiser step.  items_dev: device array of
 * n_items records {const float* w; float* img_fwd; float* img_dgrad; int Cin, Cout, k,
 * bn_fwd, bn_dgrad, passes;} (bn = min(channels,128) n-tile, passes = 2 for 3xTF32 else 1). */
typedef struct dpp_pack_item {
    const float *w;
    float *img_fwd;
    float *img_dgrad;
    int Cin, Cout, k, bn_fwd, bn_dgrad, passes;
} dpp_pack_item;
int dpp_conv_pack_all(const void *items_dev, int n_items, void *stream);

int dpp_conv2d_fwd(const dpp_conv_desc *d, const float *x, const dpp_bn_ref *in_bn,
                   const float *w, const float *bias, const float *residual, float *y,
                   double *out_stats, void *stream);

/* Backward data.  dy [N,Ho,Wo,Cout] -> dx [N,H,W,Cin].
 * accumulate != 0: dx += (used for the second consumer of a tensor).
 * If mask_bn != NULL, x_pre (the tensor the forward prologue normalised) is given and the
 * epilogue computes dz = dx * [bn(x_pre) > 0] (ReLU backward), writes dz instead, and
 * accumulates dz_stats = {sum dz, sum dz*xhat} (fp64 [2*Cin]) for the BN backward
 * (Theano's T.grad through mean/var, trainer/poseregnettrainer.py:111).
 * For stride 2 (1x1 only) dx must be zero-filled by the caller; odd positions stay 0.   */
int dpp_conv2d_dgrad(const dpp_conv_desc *d, const float *dy, const float *w, float *dx,
                     int accumulate, const dpp_bn_ref *mask_bn, const float *x_pre,
                     double *dz_stats, void *stream);

/* Backward data of the LAST consumer of a normalised tensor with that BatchNorm's backward apply fused in behind a
 * grid-wide barrier (one launch instead of dpp_conv2d_dgrad + dpp_bn_bwd_apply; same arithmetic):
 *   dz = (dz +) dgrad(dy) * [bn(x_pre) > 0], dz_stats += {sum dz, sum dz*xhat}, then
 *   dx = gamma*inv_std*(dz - mean(dz) - xhat*mean(dz*xhat)) [+ skip], dgamma/dbeta += param_grad_scale * sums.
 * gbar: two zeroed 32-bit words (device) private to this call within a step (arrival counter, error mark).
 * Returns DPP_ENOTSUP when the tcgen05 kernel does not cover the layer (stride 2, precision 0, unsupported widths,
 * deterministic statistics): the caller then issues the two separate calls.                                    */
int dpp_conv2d_dgrad_bn_bwd(const dpp_conv_desc *d, const float *dy, const float *w, float *dz,
                            int accumulate, const dpp_bn_ref *mask_bn, const float *x_pre,
                            double *dz_stats, const float *skip, float *dx, float *dgamma,
                            float *dbeta, float param_grad_scale, unsigned int *gbar, void *stream);

/* Backward weights: dw [(k*k*Cin)][Cout] += a^T dy with a = in_bn(x) recomputed on the
 * fly; db [Cout] += sum_p dy.  dw/db must be zeroed by the caller at step start.        */
int dpp_conv2d_wgrad(const dpp_conv_desc *d, const float *x, const dpp_bn_ref *in_bn,
                     const float *dy, float *dw, float *db, void *stream);

/* Grouped backward weights: the dpp_conv2d_wgrad work of MANY layers in a few persistent launches (one per n-tile
 * width 16 / 32 / 64 / 128).  A training step has 63 of these GEMMs, most too small to amortise their own launch,
 * set-up and split-reduction cost; their results are only needed by the optimiser, so they can run together once all
 * dy tensors exist.  create() copies the layer table to the device (allocates: call it outside graph capture; every
 * pointer - activations, gradients, statistics, dw / db - must stay valid while the handle lives); run() is
 * asynchronous and capturable.  dw / db are ACCUMULATED with floating-point reductions (zero them at step start).
 * create() returns DPP_ENOTSUP if a layer is outside the tcgen05 path (precision 0, unsupported widths).          */
typedef struct dpp_wgrad_layer {
    dpp_conv_desc d;
    const float *x;       /* input activations of the layer (before its BN + ReLU prologue) */
    dpp_bn_ref in_bn;     /* prologue, used when has_in_bn != 0 */
    int has_in_bn;
    const float *dy;      /* gradient w.r.t. the layer's output */
    float *dw;            /* [(k*k*Cin)][Cout] */
    float *db;            /* [Cout] or NULL */
} dpp_wgrad_layer;
int dpp_wgrad_group_create(const dpp_wgrad_layer *layers, int n_layers, void **handle_out);
int dpp_wgrad_group_run(void *handle, void *stream);
int dpp_wgrad_group_launches(void *handle);     /* kernels one run() launches */
int dpp_wgrad_group_destroy(void *handle);

/* ---- BatchNorm backward apply (full BN backward through batch statistics) -------------
 * dx = gamma*inv_std*(dz - mean(dz) - xhat*mean(dz*xhat)) [+ skip]; dgamma += sum dz*xhat,
 * dbeta += sum dz.  `bn` must be in train mode (sums != NULL).  dz and dx may alias.
 * skip (may be NULL) is the gradient arriving over the identity connection.
 * dbias_stats (may be NULL, fp64 [C]) accumulates sum_p dx for the bias gradient of the
 * conv layers that produced x (their db is that sum).
 * param_grad_scale multiplies the two parameter gradients: 1, or 1/world when dz_stats holds
 * sums over ALL data-parallel ranks (SyncBN) and the gradient arena is summed over the ranks
 * afterwards.                                                                              */
int dpp_bn_bwd_apply(const float *dz, const float *x, const dpp_bn_ref *bn,
                     const double *dz_stats, const float *skip, float *dx, float *dgamma,
                     float *dbeta, double *dbias_stats, int64_t pixels, int C,
                     float param_grad_scale, void *stream);

/* ---- SyncBN statistics exchange over peer memory (NVLink P2P) -------------------------------------------------
 * Data-parallel training with BatchNorm statistics of the GLOBAL minibatch (what the single-device reference
 * computes, net/batchnormlayer.py:154-159): the 2*C fp64 sums of every BN are summed over the ranks, forward and
 * backward.  dpp_stats_exchange is a one-shot all-reduce written for this size (a few KB, latency-bound, on the
 * critical chain): one single-CTA kernel stores the local sums into every rank's exchange buffer over NVLink,
 * publishes a sequence number, waits for the other ranks' numbers and overwrites `stats` with the sum taken in rank
 * order (bit-identical on all ranks).  Asynchronous, capturable; every rank must issue the same sequence of calls.
 *   dpp_peer_alloc / dpp_peer_open: a zeroed device buffer with its 64-byte CUDA IPC handle / the mapping of another
 *   rank's buffer (same node).  Exchange-buffer layout per sync point at `region_off` (16-byte aligned):
 *   [world][n] doubles, then [world] 64-bit flags - i.e. world * (n + 1) * 8 bytes.
 *   peers_dev: DEVICE array of `world` pointers (entry `rank` = the local buffer).  seq_counter: device u64, zero at
 *   start, advanced by the kernel.  err_flag: device u32 set to 0xDEAD if a wait gave up (a rank died).          */
int dpp_peer_alloc(int64_t bytes, void **ptr_out, unsigned char *ipc_handle_out64);
int dpp_peer_open(const unsigned char *ipc_handle64, void **ptr_out);
int dpp_peer_close(void *ptr);
int dpp_peer_free(void *ptr);
int dpp_stats_exchange(double *stats, int n, void *const *peers_dev, int64_t region_off, int rank,
                       int world, unsigned long long *seq_counter, unsigned int *err_flag, void *stream);

/* materialise a = bn(x) (+relu): used for the last BN+ReLU in front of the FC stack      */
int dpp_bn_apply(const float *x, const dpp_bn_ref *bn, float *y, int64_t pixels, int C,
                 void *stream);
/* dz = dy * [bn(x) > 0]; dz_stats += {sum dz, sum dz*xhat}                               */
int dpp_bn_relu_bwd_reduce(const float *dy, const float *x, const dpp_bn_ref *bn, float *dz,
                           double *dz_stats, int64_t pixels, int C, void *stream);
/* running-stat EMA for `count` BN layers at once (net/batchnormlayer.py:164-172):
 * mean = (1-a) mean + a batch_mean ; inv_std = (1-a) inv_std + a batch_inv_std.
 * tab_host... all arrays are device arrays of length n_layers.                           */
typedef struct dpp_bn_ema_item {
    const double *sums;
    float *mean;
    float *inv_std;
    double count;
    int C;
    float eps;
} dpp_bn_ema_item;
int dpp_bn_ema_update(const dpp_bn_ema_item *items, int n_layers, float alpha, void *stream);

/* ---- HiddenLayer (net/hiddenlayer.py:136-154): y = act(x W + b) ----------------------
 * x [B, n_in], W [n_in, n_out], y [B, n_out].  y must be zero-filled by the caller when
 * splitk > 1 ... handled internally (the call zero-fills).  mask (may be NULL): dropout
 * mask multiplied after the activation (net/dropoutlayer.py:104); scale_out multiplies the
 * result (0.7 in deterministic mode for a following DropoutLayer, else 1).              */
int dpp_fc_fwd(const float *x, const float *w, const float *bias, float *y, int B, int n_in,
               int n_out, int relu, const float *mask, float scale_out, int precision,
               void *stream);
/* dy is grad wrt the layer output (after act/mask); y is that output.  Computes
 * dpre = dy * mask * [y>0], dW += x^T dpre, db += sum dpre, dx = dpre W^T (dx may be NULL).
 * scratch must hold B*n_out floats.                                                      */
int dpp_fc_bwd(const float *x, const float *w, const float *y, const float *dy, float *dw,
               float *db, float *dx, float *scratch, int B, int n_in, int n_out, int relu,
               const float *mask, float scale_out, int precision, void *stream);
/* Same with flags.  DPP_FC_DW_ASSIGN: dW = x^T dpre (plain stores: one 67 MB write for the ResNet's first HiddenLayer
 * instead of a read-modify-write; the caller need not zero dw) - T.grad of a weight used once is an assignment
 * (trainer/poseregnettrainer.py:110-111).  db is accumulated either way.                                        */
#define DPP_FC_DW_ASSIGN 1
int dpp_fc_bwd_ex(const float *x, const float *w, const float *y, const float *dy, float *dw,
                  float *db, float *dx, float *scratch, int B, int n_in, int n_out, int relu,
                  const float *mask, float scale_out, int precision, int flags, void *stream);
/* Allocates the library-owned workspace of the k-split HiddenLayer GEMMs (9.7 MB, idempotent).  Call once outside
 * CUDA-graph capture (an uncaptured dpp_fc_fwd / dpp_fc_bwd allocates it lazily; a captured one without it reduces
 * with red.global.add instead).                                                                                 */
int dpp_fc_workspace_init(void);

/* ---- cost (trainer/poseregnettrainer.py:84-99) ----------------------------------------
 * out, target [B, D]; cost = mean_b sum_d (out-target)^2 (numJoints==1 branch) written to
 * cost_out[0] (device); dout = 2 (out-target)/B.                                          */
int dpp_loss_sqerr(const float *out, const float *target, float *dout, float *cost_out, int B,
                   int D, void *stream);

/* ---- ADAM (trainer/optimizer.py:58-90) over one flat parameter arena -------------------
 * hyper (device, 4 floats): {lr, t, unused, grad_scale}; t is incremented by the kernel
 * launched with n == 0 ... see dpp_adam_tick.  w, g, m, v are flat arrays of n floats.   */
int dpp_adam_step(float *w, const float *g, float *m, float *v, const float *hyper,
                  int64_t n, void *stream);
int dpp_adam_tick(float *hyper, void *stream); /* t += 1 after all dpp_adam_step calls    */

/* fill helpers (capturable) */
/* Strided device-to-device copy of `rows` rows of `width` bytes (pitches in bytes): concatenation of the
 * flattened ScaleNet tower outputs (reference: net/scalenet.py:174-178). */
int dpp_copy2d(void *dst, int64_t dst_pitch, const void *src, int64_t src_pitch, int64_t width, int64_t rows,
               void *stream);
int dpp_fill_f32(float *p, float value, int64_t n, void *stream);
int dpp_fill_f64(double *p, double value, int64_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DPP_B200_H */
