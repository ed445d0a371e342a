/* dcgansr.h -- C ABI of libdcgansr.so, the B200-native (sm_100a) DCGAN super-resolution
 * training step.
 *
 * This header is the drop-in boundary.  The reference
 * (PJunhyuk/dcgan-super-resolution) has no FFI of its own: its seam is the Lua method
 * surface of Torch7 objects used by train*.lua.  Every entry point below cites the
 * reference interface (file:line under /root/reference) it replaces.  The header is a
 * C99 subset with no macros in signatures so LuaJIT `ffi.cdef`, Python `cffi`/`ctypes`
 * and a C compiler all parse the same text (lua/dcgansr.lua loads it verbatim).
 *
 * Conventions
 *   - every call returns int: 0 = DCGANSR_OK, <0 = error (dcgansr_last_error gives text);
 *     no C++ exception crosses the boundary;
 *   - host tensors are caller-owned fp32, NCHW contiguous (Torch7 FloatTensor layout);
 *     weights use the Torch7 layouts (conv [nOut][nIn][kH][kW], full-conv
 *     [nIn][nOut][kH][kW], BN gamma then beta) and the flat order of
 *     Module:getParameters() (train.lua:202-203);
 *   - the library owns all device memory; internally activations are NHWC;
 *   - one host thread per ctx; distinct ctxs are independent;
 *   - there is NO CPU fallback: every compute entry point needs a CUDA device.
 */
#ifndef DCGANSR_H
#define DCGANSR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* FFI-CDEF-BEGIN  (lua/dcgansr.lua feeds everything up to FFI-CDEF-END to LuaJIT ffi.cdef verbatim) */
typedef struct dcgansr_ctx dcgansr_ctx;
typedef struct dcgansr_net dcgansr_net;

/* status codes */
enum {
  DCGANSR_OK = 0,
  DCGANSR_ERR_INVALID = -1,     /* bad argument / shape */
  DCGANSR_ERR_CUDA = -2,        /* CUDA runtime / driver error */
  DCGANSR_ERR_NOMEM = -3,
  DCGANSR_ERR_UNSUPPORTED = -4,
  DCGANSR_ERR_NCCL = -5
};

/* precision modes (north_star: strict fp32 <= 1e-5, fast TF32 <= 2e-3) */
enum {
  DCGANSR_STRICT_FP32 = 0,      /* FFMA SIMT convolutions, fp32 everywhere */
  DCGANSR_FAST_TF32 = 1         /* tcgen05 kind::tf32 implicit-GEMM convolutions, fp32 storage/accumulate */
};

/* layer kinds: 1:1 with the netG:add(...) / netD:add(...) calls (train.lua:97-136) */
enum {
  DCGANSR_CONV = 1,             /* nn.SpatialConvolution          train.lua:108 */
  DCGANSR_FULLCONV = 2,         /* nn.SpatialFullConvolution      train.lua:99 */
  DCGANSR_BN = 3,               /* nn.SpatialBatchNormalization   train.lua:100 */
  DCGANSR_RELU = 4,             /* nn.ReLU(true)                  train.lua:100 */
  DCGANSR_LRELU = 5,            /* nn.LeakyReLU(0.2,true)         train.lua:109 */
  DCGANSR_TANH = 6,             /* nn.Tanh                        train.lua:112 */
  DCGANSR_SIGMOID = 7,          /* nn.Sigmoid                     train.lua:134 */
  DCGANSR_UPNEAREST = 8,        /* nn.SpatialUpSamplingNearest    train-gray.lua:104 */
  DCGANSR_VIEW = 9              /* nn.View(1):setNumInputDims(3)  train.lua:136 */
};

/* activation kinds for the layer-level ops */
enum { DCGANSR_ACT_NONE = 0, DCGANSR_ACT_RELU = 1, DCGANSR_ACT_LRELU = 2, DCGANSR_ACT_TANH = 3, DCGANSR_ACT_SIGMOID = 4 };

/* loss families */
enum {
  DCGANSR_LOSS_BCE = 0,         /* nn.BCECriterion   train-gray-patch.lua:113 */
  DCGANSR_LOSS_MSE = 1          /* nn.MSECriterion   train.lua:142 */
};

typedef struct dcgansr_cfg {
  int device;                   /* CUDA ordinal (cutorch.setDevice, train.lua:169; 0-based here) */
  int precision;                /* DCGANSR_STRICT_FP32 | DCGANSR_FAST_TF32 */
  int world_size;               /* data-parallel ranks (1 = single GPU) */
  int rank;
  int sync_bn;                  /* 1: BN batch statistics all-reduced across ranks (exact big-batch semantics) */
  int use_graph;                /* 1: capture dcgansr_train_step into a CUDA graph and replay it */
} dcgansr_cfg;

typedef struct dcgansr_layer {
  int kind;
  int cin, cout;                /* conv / fullconv; BN uses cout = channels */
  int kh, kw, sh, sw, ph, pw, adjh, adjw;
  float negval;                 /* LeakyReLU slope */
  float eps, momentum;          /* BN */
  int scale;                    /* UPNEAREST factor */
} dcgansr_layer;

typedef struct dcgansr_step_cfg {
  int loss;                     /* DCGANSR_LOSS_BCE | DCGANSR_LOSS_MSE */
  float real_label;             /* label:fill for D(real)      train.lua:219 */
  float fake_label;             /* label:fill for D(fake)      train-gray-patch.lua:303 */
  float gen_label;              /* label:fill in fGx           train.lua:264 */
  int pixel_label;              /* 1: D(fake) target = per-sample pixel MSE (label:copy(errVal_PSNR), train.lua:245) */
  float pixel_div;              /* divisor of the per-sample squared error sum (train.lua:194: 4*C*H*W) */
  double lr, beta1, beta2, eps; /* optim.adam config (train.lua:145-152; defaults of optim/adam.lua); Lua numbers are doubles */
} dcgansr_step_cfg;

/* ---- lifecycle ------------------------------------------------------------------------- */
int dcgansr_version(void);
/* replaces `require 'cunn'; cutorch.setDevice(1)` (train.lua:168-169) */
int dcgansr_ctx_create(const dcgansr_cfg* cfg, dcgansr_ctx** out);
void dcgansr_ctx_destroy(dcgansr_ctx* ctx);
const char* dcgansr_last_error(dcgansr_ctx* ctx);   /* ctx may be NULL: last error of the calling thread */
int dcgansr_synchronize(dcgansr_ctx* ctx);
/* CUDA-event timers on the library's stream (replaces torch.Timer, train.lua:159-161) */
int dcgansr_timer_begin(dcgansr_ctx* ctx);
int dcgansr_timer_end(dcgansr_ctx* ctx, float* ms_out);
/* number of kernel launches issued by this ctx since creation (graph replays count their nodes) */
int dcgansr_launch_count(dcgansr_ctx* ctx, int64_t* out);
/* writes a buffer larger than L2 on the library's stream (bench hygiene) */
int dcgansr_flush_l2(dcgansr_ctx* ctx);
/* per-launch CUDA-event profiler on the library's stream (new; the reference only has torch.Timer):
 * between begin and end every kernel launch is bracketed by events.  end writes a JSON array
 * [{"name","work","kind":"flops"|"bytes","launches","ms"}...] sorted by time; `work` is the
 * ALGORITHMIC flops/bytes of ONE launch.  Not usable while a CUDA graph replays the step. */
int dcgansr_profile_begin(dcgansr_ctx* ctx);
int dcgansr_profile_end(dcgansr_ctx* ctx, char* json_out, int64_t cap);

/* ---- data-parallel communicator (new design; no reference counterpart) -- */
/* NCCL is dlopen'ed lazily.  unique_id is the 128-byte ncclUniqueId, produced on rank 0 by
 * dcgansr_comm_get_unique_id and distributed by the host (torch.distributed / MPI / file). */
int dcgansr_comm_get_unique_id(dcgansr_ctx* ctx, void* unique_id_128);
int dcgansr_comm_init(dcgansr_ctx* ctx, const void* unique_id_128);
/* 1 when dcgansr_comm_init mapped every rank's exchange area into every peer (cudaIpc over NVLink / NVSwitch, one node,
 * <= 8 ranks): the sync_bn statistics are then all-reduced by a one-shot push kernel fused into the BatchNorm statistics tail
 * instead of ncclAllReduce.  0: single rank, peers not mappable, or DCGANSR_PEER_AR=0 -- every rank takes the same answer. */
int dcgansr_comm_peer_enabled(dcgansr_ctx* ctx);

/* ---- net description: nn.Sequential():add(...) (train.lua:97-136) ------------------------ */
/* ctx may be NULL: a plan-only net (shape inference, parameter counts; no device memory, every
 * compute entry point then returns DCGANSR_ERR_INVALID). */
int dcgansr_net_create(dcgansr_ctx* ctx, const dcgansr_layer* layers, int n_layers,
                       int in_c, int in_h, int in_w, int max_batch, dcgansr_net** out);
void dcgansr_net_destroy(dcgansr_net* net);
int dcgansr_net_out_shape(dcgansr_net* net, int* c, int* h, int* w);
/* Module:getParameters() (train.lua:202-203): flat fp32 vectors in module order */
int dcgansr_net_num_params(dcgansr_net* net, int64_t* out);
int dcgansr_net_set_params(dcgansr_net* net, const float* host_flat);
int dcgansr_net_get_params(dcgansr_net* net, float* host_flat);
int dcgansr_net_get_grads(dcgansr_net* net, float* host_flat);
/* BN running_mean / running_var of all BN modules, concatenated in module order */
int dcgansr_net_num_bn_channels(dcgansr_net* net, int64_t* out);
int dcgansr_net_get_bn_running(dcgansr_net* net, float* mean, float* var);
int dcgansr_net_set_bn_running(dcgansr_net* net, const float* mean, const float* var);
/* optimState{m,v,t} (train.lua:145-152; fields written by optim/adam.lua) */
int dcgansr_net_get_adam_state(dcgansr_net* net, float* m, float* v, int64_t* t);
int dcgansr_net_set_adam_state(dcgansr_net* net, const float* m, const float* v, int64_t t);

/* ---- Torch7-shaped net ops (host NCHW fp32 in/out; H2D/D2H inside) ----------------------- */
/* net:forward(x) (train.lua:218).  y may be NULL.  Caches activations: last forward wins. */
int dcgansr_net_forward(dcgansr_net* net, const float* x, int batch, float* y);
/* net:backward(x, dy) (train.lua:222): dgrad + accumulate wgrad.  dx may be NULL. */
int dcgansr_net_backward(dcgansr_net* net, const float* x, const float* dy, int batch, float* dx);
/* net:updateGradInput(x, dy) (train.lua:268): dgrad only, current weights, cached activations. */
int dcgansr_net_update_grad_input(dcgansr_net* net, const float* x, const float* dy, int batch, float* dx);
/* gradParameters:zero() (train.lua:209) */
int dcgansr_net_zero_grads(dcgansr_net* net);
/* optim.adam(feval, parameters, optimState) (train.lua:280) on the net's flat params/grads */
int dcgansr_net_adam(dcgansr_net* net, double lr, double beta1, double beta2, double eps);

/* ---- the fused step: fDx -> adam(D) -> fGx -> adam(G) (train.lua:208-283) ----------------- */
/* real_host: this rank's shard of the minibatch, NCHW fp32, local_batch samples of D's input
 * shape.  out_losses (may be NULL: no host sync) receives errD_real, errD_fake, errG averaged
 * over the global batch. */
int dcgansr_train_step(dcgansr_ctx* ctx, dcgansr_net* netG, dcgansr_net* netD,
                       const dcgansr_step_cfg* cfg, const float* real_host, int local_batch,
                       float* out_losses);
/* Same step with the batch already resident on the device (staged by dcgansr_stage_batch);
 * this is the "inputs resident in HBM" leg of the benchmark. */
int dcgansr_stage_batch(dcgansr_ctx* ctx, dcgansr_net* netD, const float* real_host, int local_batch, int slot);
int dcgansr_train_step_staged(dcgansr_ctx* ctx, dcgansr_net* netG, dcgansr_net* netD,
                              const dcgansr_step_cfg* cfg, int slot, int local_batch, float* out_losses);
/* Patch extraction / re-assembly on the device (train-gray-patch.lua:267-273,588-595; train-gray-patch-batch.lua:258-264,
 * 434-442; train-gray-patch-batch-overlap.lua:393-399).  Single-channel images [k][h][w]; patch i of image j:
 *   patches[j*nper + i][a][b] = images[j][(i / line) * stride + a][(i % line) * stride + b]
 * (line, stride) = (patchSize, patchSize) in the non-overlapping scripts, (overlapPatchLine, overlap) in the overlapping one.
 * assemble is the inverse scatter (where patches overlap the highest patch index wins, as in the reference loop; pixels
 * no patch covers keep the value passed in).  stage_patches makes the patches the staged batch of `slot`. */
int dcgansr_extract_patches(dcgansr_ctx* ctx, const float* images, float* patches, int k, int h, int w, int patch,
                            int line, int nper, int stride);
int dcgansr_assemble_patches(dcgansr_ctx* ctx, const float* patches, float* images, int k, int h, int w, int patch,
                             int line, int nper, int stride);
int dcgansr_stage_patches(dcgansr_ctx* ctx, dcgansr_net* netD, const float* images_host, int k, int h, int w, int patch,
                          int line, int nper, int stride, int slot);
/* Overlap stitching (train-gray-patch-batch-overlap.lua:457-694): k images' worth of generated patches
 * [k * line * line][patch][patch], line = (h - overlap) / (patch - overlap) (:387), patch i at (i / line, i % line) * overlap,
 * are joined along minimum-error boundary cuts (dynamic programme over |neighbour - current| on each overlap strip, float64
 * cost tables as in the reference) into images [k][h][w]; pixels no patch covers keep the value passed in (the reference
 * starts from zeros).  The patch order, tie-breaking and the fact that an interior patch's left seam overwrites its top
 * seam follow the reference.  flags bit 0: take the top-seam cost against the patch above (i - line) instead of the
 * reference's patch i - 1 (:557). */
int dcgansr_stitch_overlap(dcgansr_ctx* ctx, const float* patches, float* images, int k, int h, int w, int patch,
                           int overlap, int flags);
/* Evaluation metrics of the eval sweeps on n single-channel h x w image pairs (host pointers, one value per pair):
 * calPSNR (train-gray-3.lua:143-151): 10*log10(1/MSE), MSE = sum((a-b)^2)/(h*w), 99 when MSE == 0;
 * calSSIM (train-gray-3.lua:156-221): images in [-1,1] mapped to [0,255], 11x11 Gaussian (sigma 1.5) 'full' convolution,
 * K1 = 0.01, K2 = 0.03, L = 255, mean of the SSIM map. */
int dcgansr_psnr(dcgansr_ctx* ctx, const float* a, const float* b, float* out, int n, int h, int w);
int dcgansr_ssim(dcgansr_ctx* ctx, const float* a, const float* b, float* out, int n, int h, int w);
/* The sweeps' bilinear baseline: image.scale(src, dw, dh) in its default 'bilinear' mode (train-gray-3.lua:399) on n
 * single-channel h x w images -> dh x dw.  Separable linear interpolation, end points aligned, rows first, float32
 * intermediate (upstream torch/image).  Only enlarging sizes (dh >= h, dw >= w). */
int dcgansr_scale_bilinear(dcgansr_ctx* ctx, const float* src, float* dst, int n, int h, int w, int dh, int dw);
/* netG:forward on a batch of low-res inputs (eval path, train-gray-3.lua:359-445) */
int dcgansr_generate(dcgansr_ctx* ctx, dcgansr_net* netG, const float* lr_host, int batch, float* sr_host);

/* ---- layer-level ops (parity tests; host NCHW fp32; stateless) ---------------------------- */
/* nn.SpatialConvolution forward / updateGradInput / accGradParameters (train.lua:108) */
int dcgansr_conv2d_fwd(dcgansr_ctx* ctx, const float* x, const float* w, float* y,
                       int n, int cin, int h, int wd, int cout, int k, int s, int p);
int dcgansr_conv2d_dgrad(dcgansr_ctx* ctx, const float* dy, const float* w, float* dx,
                         int n, int cin, int h, int wd, int cout, int k, int s, int p);
int dcgansr_conv2d_wgrad(dcgansr_ctx* ctx, const float* x, const float* dy, float* dw,
                         int n, int cin, int h, int wd, int cout, int k, int s, int p);
/* nn.SpatialFullConvolution (train.lua:99); h, wd are the INPUT spatial dims */
int dcgansr_fullconv2d_fwd(dcgansr_ctx* ctx, const float* x, const float* w, float* y,
                           int n, int cin, int h, int wd, int cout, int k, int s, int p);
int dcgansr_fullconv2d_dgrad(dcgansr_ctx* ctx, const float* dy, const float* w, float* dx,
                             int n, int cin, int h, int wd, int cout, int k, int s, int p);
int dcgansr_fullconv2d_wgrad(dcgansr_ctx* ctx, const float* x, const float* dy, float* dw,
                             int n, int cin, int h, int wd, int cout, int k, int s, int p);
/* Measurement aid (no reference counterpart): average device time in ms of one conv / full-conv op
 * (what: 0 forward, 1 updateGradInput, 2 accGradParameters) over `iters` launches on device-resident tensors. */
int dcgansr_bench_conv(dcgansr_ctx* ctx, int full, int what, int n, int cin, int h, int wd, int cout,
                       int k, int s, int p, int iters, float* ms_out);

/* nn.SpatialBatchNormalization training forward/backward (train.lua:100) */
int dcgansr_bn_fwd_train(dcgansr_ctx* ctx, const float* x, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, float* y,
                         float* save_mean, float* save_invstd,
                         int n, int c, int h, int wd, float eps, float momentum);
int dcgansr_bn_bwd(dcgansr_ctx* ctx, const float* x, const float* dy, const float* gamma,
                   const float* save_mean, const float* save_invstd,
                   float* dx, float* dgamma, float* dbeta, int n, int c, int h, int wd);
/* nn.ReLU / LeakyReLU / Tanh / Sigmoid; backward takes the activation OUTPUT y */
int dcgansr_act_fwd(dcgansr_ctx* ctx, const float* x, float* y, int64_t count, int kind, float negval);
int dcgansr_act_bwd(dcgansr_ctx* ctx, const float* y, const float* dy, float* dx, int64_t count, int kind, float negval);
/* nn.SpatialUpSamplingNearest(2) (train-gray.lua:104) */
int dcgansr_upnearest2_fwd(dcgansr_ctx* ctx, const float* x, float* y, int n, int c, int h, int wd);
int dcgansr_upnearest2_bwd(dcgansr_ctx* ctx, const float* dy, float* dx, int n, int c, int h, int wd);
/* the 2x2 box down-sample loop (train.lua:225-230) */
int dcgansr_avgpool2_fwd(dcgansr_ctx* ctx, const float* x, float* y, int n, int c, int h, int wd);
/* criterion:forward / :backward (train.lua:220-221); label has `count` entries */
int dcgansr_bce(dcgansr_ctx* ctx, const float* x, const float* label, int64_t count, float* loss, float* dx);
int dcgansr_mse(dcgansr_ctx* ctx, const float* x, const float* label, int64_t count, float* loss, float* dx);
/* calMSE loop (train.lua:193-195,237-239): out[b] = sum((real-fake)^2)/div */
int dcgansr_pixel_mse_per_sample(dcgansr_ctx* ctx, const float* real, const float* fake, float* out,
                                 int n, int64_t per_sample, float div);
/* optim.adam on caller vectors (train.lua:280); t is the step count BEFORE the update */
int dcgansr_adam_step(dcgansr_ctx* ctx, float* p, const float* g, float* m, float* v, int64_t count,
                      int64_t t, double lr, double beta1, double beta2, double eps);
/* FFI-CDEF-END */

#ifdef __cplusplus
}
#endif
#endif /* DCGANSR_H */
