/* fsr_b200.h - C ABI of libfsr_b200.so: B200 (sm_100a) kernels for the Fast-SRGAN hot path.
 *
 * The reference (HasnainRaz/Fast-SRGAN) has NO native/FFI interface: its hot path is the Python
 * class surface of model.py / trainer.py executed by ATen.  Each entry point below therefore cites
 * the reference torch.nn call site(s) (file:line) whose arithmetic it replaces.  The Python mirror
 * (fast-srgan_b200/model.py, trainer.py) binds these with ctypes - see INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types.  All pointers are DEVICE pointers unless noted.
 *   - the caller owns every buffer (activations, packed weights, workspace); nothing is allocated
 *     or freed here.  Launches go on `stream` (a cudaStream_t passed as void*); calls are
 *     asynchronous and re-entrant per stream.
 *   - return 0 on success; <0 on error (fsr_error_string()).  -(1000+e) wraps cudaError_t e.
 *   - activations between kernels are NHWC, `dtype` FSR_F16 or FSR_BF16 (fp32 accumulation
 *     everywhere); the module boundary (neck input / head output) is the reference's fp32 NCHW
 *     (model.py:112-117) or inference.py's uint8 HWC (inference.py:48-56).
 */
#ifndef FSR_B200_H_
#define FSR_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSR_ABI_VERSION 2
#define FSR_MAX_LAYERS 32

enum { FSR_F16 = 0, FSR_BF16 = 1 };
/* conv epilogues (fused neighbours of the conv in the reference graph) */
enum {
  FSR_EPI_RAW_STATS = 0, /* raw conv out + InstanceNorm sum/sumsq   (model.py:54-55,64-65,93-94,131-132) */
  FSR_EPI_BIAS_ACT = 1,  /* act(conv + bias)                         (VGG conv+ReLU, model.py:8)         */
  FSR_EPI_PS_PRELU = 2,  /* bias + PixelShuffle(2) + PReLU           (model.py:39-40)                    */
  FSR_EPI_HEAD_TANH = 3, /* bias + tanh -> fp32 NCHW | uint8 NHWC    (model.py:102-110, inference.py:54-56) */
  FSR_EPI_F32 = 4        /* precise mode: out fp32 NHWC [N,H,W,cout] = conv (act = 0) or out + conv (act = 1); see below */
};
enum { FSR_ACT_NONE = 0, FSR_ACT_RELU = 1, FSR_ACT_LRELU = 2, FSR_ACT_PRELU = 3 };

int fsr_abi_version(void);
const char* fsr_error_string(int code);

/* Pack torch OIHW fp32 conv weights [cout,cin,3,3] into the kernel layout [9][cout_pad][cin] (dtype).
 * ps_perm=1 additionally permutes output channels so that PixelShuffle(2) (model.py:36) becomes a
 * contiguous store: packed row (2i+j)*(cout/4)+c <- reference channel 4c+2i+j.  bias (nullable) is
 * permuted/padded the same way into bias_packed[cout_pad] (fp32). */
int fsr_pack_conv3x3_weight(const float* w_oihw, const float* bias, void* w_packed, float* bias_packed, int cout,
                            int cin, int cout_pad, int ps_perm, int dtype, void* stream);

/* Up to 48 packs (any mix of fsr_pack_conv3x3_weight and fsr_pack_conv3x3_weight_t) in ONE launch: the training step
 * re-packs every conv of a network after each AdamW step (trainer.py:181,196).  `tasks` is a HOST array; pad = cout_pad
 * (forward pack) or row_pad (transposed pack); flags: FSR_PACK_T = transposed (data-gradient) pack, FSR_PACK_PS = ps_perm,
 * FSR_PACK_FLIP = flipped taps; bias / bias_out / row_scale nullable. */
#define FSR_PACK_T 1
#define FSR_PACK_PS 2
#define FSR_PACK_FLIP 4
typedef struct {
  const float* w; void* out; const float* bias; float* bias_out; const float* row_scale;
  int cout, cin, pad, flags;
} FsrPackTask;
int fsr_pack_multi(const FsrPackTask* tasks, int n, int dtype, void* stream);

/* 3x3 / stride 1 / pad 1 convolution, Cin = 64, on tcgen05 tensor cores (implicit GEMM, TMA-fed).
 * Replaces torch.nn.Conv2d at model.py:30-35 (UpSamplingBlock.conv), :47-54/:57-64 (ResidualBlock
 * conv1/conv2), :87-93 (bottleneck), :103-108 (head).
 *   x        [N,H,W,64] NHWC dtype        w_packed  from fsr_pack_conv3x3_weight
 *   epilogue FSR_EPI_*:
 *     RAW_STATS : out [N,H,W,cout] dtype, stats [N,cout,2] int64 FIXED POINT += (sum * 2^24, sumsq * 2^20); caller zeroes
 *                 (integer atomics are order independent: the forward pass is bitwise reproducible)
 *     BIAS_ACT  : out [N,H,W,cout] dtype = act(conv + bias)
 *     PS_PRELU  : cout = 256: out [N,2H,2W,64] dtype = PReLU(PixelShuffle2(conv + bias)); alpha = device ptr
 *     HEAD_TANH : cout_pad = 16 (3 real): out fp32 [N,3,H,W] (out_u8=0) or uint8 [N,H,W,3] (out_u8=1);
 *                 out_u8=2|3: no tanh, fp32 NCHW store | accumulate (the 64->3 data gradients of the
 *                 3-channel first layers: VGG conv1_1, Discriminator.neck)
 */
int fsr_conv3x3_c64(const void* x, const void* w_packed, void* out, const float* bias, int64_t* stats,
                    const float* alpha, int N, int H, int W, int cout, int epilogue, int act, float slope,
                    int out_u8, int dtype, void* stream);

/* General 3x3 / pad 1 convolution on tcgen05: cin, cout multiples of 64, stride 1 or 2, forward (mode 0)
 * or data gradient (mode 1).  Replaces torch.nn.Conv2d at model.py:124-131 (SimpleBlock.conv), the
 * torchvision VGG19 convs behind model.py:8, and autograd's convolution_backward (input gradient) for
 * trainer.py:180,195.  H, W = spatial size of the conv INPUT (forward) / of dX (mode 1).
 *   mode 0, stride 1: x [N,H,W,cin] NHWC -> out [N,H,W,cout]
 *   mode 0, stride 2: x in parity-plane layout [N][4][H/2][W/2][cin] -> out [N,H/2,W/2,cout] NHWC
 *   mode 1, stride 1: x = dY [N,H,W,cin=Cout_fwd]; w packed with transpose=1 -> out = dX [N,H,W,cout=Cin_fwd]
 *   mode 1, stride 2: x = dY [N,H/2,W/2,cin]; out = dX in parity-plane layout [N][4][H/2][W/2][cout]
 * epilogue: FSR_EPI_RAW_STATS (out + InstanceNorm sum/sumsq), FSR_EPI_BIAS_ACT, or FSR_EPI_PS_PRELU (mode 0, stride 1,
 * cout = 4F packed with ps_perm: bias + PReLU + PixelShuffle(2) -> out [N,2H,2W,F]; the UpSamplingBlock for F != 64). */
int fsr_conv3x3_gen(const void* x, const void* w_packed, void* out, const float* bias, int64_t* stats,
                    const float* alpha, int N, int H, int W, int cin, int cout, int stride, int mode, int epilogue,
                    int act, float slope, int dtype, void* stream);

/* Conv2d(cin -> 3, k3, p1) (+ tanh): Generator.head for any n_filters (model.py:102-110), cin multiple of 64 (<= 512),
 * w_packed [9][16][cin] (3 real rows).  out_mode: 0 tanh -> fp32 NCHW, 1 tanh -> uint8 NHWC, 2|3 linear store|accumulate. */
int fsr_conv3x3_head(const void* x, const void* w_packed, void* out, const float* bias, int N, int H, int W, int cin,
                     int out_mode, int dtype, void* stream);

/* Conv2d(3 -> cout, k3, p1) + bias + activation, direct (HBM-bound; K = 27 is no tensor-core shape).
 * Replaces model.py:75-78 (Generator.neck, PReLU) and :143-146 (Discriminator.neck, LeakyReLU 0.2);
 * vgg_norm=1 also folds VGG19.forward's renormalisation model.py:21-22 into the load (VGG conv1_1).
 *   x: fp32 NCHW [N,3,H,W] (in_u8=0) or uint8 NHWC [N,H,W,3] (in_u8=1: x/127.5-1, inference.py:50)
 *   w: fp32 OIHW [cout,3,3,3], bias fp32 [cout]; out NHWC dtype [N,H,W,cout]; cout % 64 == 0 */
int fsr_neck_conv3x3(const void* x, const float* w, const float* bias, const float* alpha, void* out, int N, int H,
                     int W, int cout, int act, float slope, int in_u8, int vgg_norm, int dtype, void* stream);

/* InstanceNorm2d(affine=False, eps) normalise from conv-epilogue statistics, fused with the following
 * activation and residual add: out = act((raw-mean)*rstd) (+ residual).
 * Replaces model.py:55-56 (bn1+relu1), :65+:69 (bn2 + skip), :94+:115 (bottleneck IN + long skip),
 * :132-133 (SimpleBlock bn + LeakyReLU).  raw/out/residual NHWC dtype [N,HW,C]; stats [N,C,2] int64 fixed point. */
int fsr_instnorm_apply(const void* raw, const int64_t* stats, const void* residual, void* out, const float* alpha,
                       int N, int HW, int C, int act, float slope, float eps, int dtype, void* stream);
/* same normalise + activation, output written in the parity-plane layout [N][4][H/2][W/2][C] that the stride-2
 * fsr_conv3x3_gen reads (SimpleBlock with stride 2, model.py:148-183): folds fsr_parity_layout into this pass.
 * H, W even; out must not alias raw. */
int fsr_instnorm_apply_parity(const void* raw, const int64_t* stats, void* out, const float* alpha, int N, int H, int W, int C,
                              int act, float slope, float eps, int dtype, void* stream);

/* torch.nn.PixelShuffle(2) (model.py:36) on NHWC: in [N,H,W,4C] (reference channel order) -> out [N,2H,2W,C]. */
int fsr_pixel_shuffle2(const void* in, void* out, int N, int H, int W, int C, int dtype, void* stream);

/* Module-boundary layout conversion (the reference is NCHW fp32 everywhere, inference.py:50-51). */
int fsr_nchw_f32_to_nhwc(const float* in, void* out, int N, int C, int HW, int dtype, void* stream);
int fsr_nhwc_to_nchw_f32(const void* in, float* out, int N, int C, int HW, int dtype, void* stream);

/* ====================== GAN training step (trainer.py:168-196): backward, losses, optimiser ======================
 * Tensors are NHWC `dtype` unless stated; gradients of parameters are fp32 in torch's OIHW layout, ACCUMULATED (+=). */

/* dgrad pack: out[tap'][ci (padded to row_pad)][col] = row_scale[ci] * W[oc(col)][ci][flip ? 8-tap' : tap']
 * (rows = forward input channel, K = forward output channel).  flip=0 feeds fsr_conv3x3_gen(mode=1) (cin := Cout_fwd,
 * cout := Cin_fwd); flip=1 feeds fsr_conv3x3_c64 (a data gradient IS a stride-1 conv with transposed, flipped
 * weights).  row_scale (nullable) folds VGG19.forward's renormalisation chain rule (model.py:21-22). */
int fsr_pack_conv3x3_weight_t(const float* w_oihw, void* w_packed, int cout, int cin, int ps_perm, int flip, int row_pad,
                              const float* row_scale, int dtype, void* stream);

/* Weight gradient of a 3x3/pad-1 conv on tcgen05 (autograd convolution_backward, weight part; trainer.py:180,195):
 * dw[co,ci,r,s] += sum dY[n,y,x,co] * X[n, stride*y+r-1, stride*x+s-1, ci].  H, W = X spatial size;
 * stride 2: X in parity-plane layout.  ps_perm: dY columns are pixel-shuffle-permuted (UpSamplingBlock).
 * Two stages without atomics (split-K partial tiles in `workspace`, then a fixed-order reduction): the result is
 * bitwise reproducible.  workspace: >= fsr_wgrad_workspace_bytes() bytes of device memory (contents are scratch). */
size_t fsr_wgrad_workspace_bytes(void);
int fsr_conv3x3_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int cin, int cout, int stride,
                      int ps_perm, void* workspace, size_t ws_bytes, int dtype, void* stream);
/* `groups` (<= 40) weight gradients of IDENTICAL shape (stride 1) in ONE launch - the generator's 2L+1 64->64 convs
 * (model.py:47-64, 87-93): the conv inputs sit in x_arena [groups][N,H,W,cin] and the output gradients in dy_arena
 * [groups][N,H,W,cout] (group strides in elements); dw_list_host[g] (HOST array of device pointers) receives group g. */
int fsr_conv3x3_wgrad_grouped(const void* x_arena, const void* dy_arena, float* const* dw_list_host, int groups,
                              long long x_group_stride, long long dy_group_stride, int N, int H, int W, int cin, int cout,
                              void* workspace, size_t ws_bytes, int dtype, void* stream);

/* NHWC [N,H,W,C] <-> parity planes [N][4][H/2][W/2][C] (input layout of stride-2 convs, model.py:124-131). */
int fsr_parity_layout(const void* in, void* out, int N, int H, int W, int C, int to_parity, int dtype, void* stream);

/* torch.nn.MaxPool2d(2) of VGG19.features (model.py:8) and its backward fused with the preceding ReLU's. */
int fsr_maxpool2(const void* in, void* out, int N, int H, int W, int C, int dtype, void* stream);
int fsr_maxpool2_relu_bwd(const void* in, const void* dout, void* din, int N, int H, int W, int C, int dtype, void* stream);
int fsr_relu_bwd(const void* y, const void* dy, void* dx, size_t n_elems, int dtype, void* stream);
int fsr_add(const void* a, const void* b, void* out, size_t n_elems, int dtype, void* stream);

/* Discriminator's final Conv2d(8F -> 1, k1) (model.py:184-186): fp32 logits z[npix]; backward. */
int fsr_conv1x1_to1_fwd(const void* x, const float* w, const float* b, float* z, int npix, int C, int dtype, void* stream);
int fsr_conv1x1_to1_bwd(const void* x, const float* w, const float* dz, void* dx, float* dw, float* db, int npix, int C,
                        int dtype, void* stream);

/* BCEWithLogitsLoss (trainer.py:41,177,178,188) with labels t = lab_scale*noise + lab_shift (trainer.py:175,176,187):
 * loss_out[0] = mean loss; dz (nullable) = grad_scale * dloss/dz. */
int fsr_bce_logits(const float* z, const float* noise, float lab_scale, float lab_shift, int n, float* loss_out, float* dz,
                   float grad_scale, void* stream);
/* SmoothL1Loss(beta=1) (trainer.py:43,192 / :109): loss_acc[0] += SUM of elementwise losses (caller divides by n);
 * da (nullable) = grad_scale * dsum/da.  dtype 2 = fp32 operands (pretrain step on images). */
int fsr_smooth_l1(const void* a, const void* b, size_t n, float* loss_acc, void* da, float grad_scale, int dtype, void* stream);

/* InstanceNorm2d (+PReLU|LeakyReLU) backward (model.py:55-56,65,94,132-133): from the conv output `raw`, its forward
 * statistics and dY, writes dRaw; PReLU slope gradient accumulated into dalpha.  red = [N][C][2] fp32 scratch. */
int fsr_instnorm_bwd(const void* raw, const int64_t* stats, const void* dy, float* red, void* draw, const float* alpha,
                     float* dalpha, int N, int HW, int C, int act, float slope, float eps, int dtype, void* stream);
/* same with dy given in the parity-plane layout [N][4][H/2][W/2][C] that a stride-2 data gradient (fsr_conv3x3_gen mode 1)
 * writes: folds the re-layout pass into the backward (planes <= 64x64, C % 16 == 0). */
int fsr_instnorm_bwd_parity(const void* raw, const int64_t* stats, const void* dy_parity, void* draw, const float* alpha, float* dalpha,
                            int N, int H, int W, int C, int act, float slope, float eps, int dtype, void* stream);
/* activation backward from the stored post-activation tensor (neck PReLU / LeakyReLU). */
int fsr_act_bwd(const void* y, const void* dy, void* dv, size_t n_elems, const float* alpha, float slope, int act,
                float* dalpha, int dtype, void* stream);
/* UpSamplingBlock (model.py:39-40) backward glue: dU [N,2H,2W,F] -> dConv [N,H,W,4F] (permuted columns). */
int fsr_ps_prelu_bwd(const void* U, const void* dU, void* dconv, int N, int H, int W, int F, const float* alpha, float* dalpha,
                     int dtype, void* stream);
/* head tanh backward (model.py:109): dpre = dy * (1 - y^2), fp32 NCHW. */
int fsr_tanh_bwd(const float* y, const float* dy, float* dpre, size_t n, void* stream);
/* weight gradient of a 3x3 conv with a 3-channel side (necks: flip=0, img = conv input; head: flip=1, img = dpre):
 * out[(c3*9 + tap)*C64 + c64] += sum img[n,c3,y+dy,x+dx] * act[n,y,x,c64]. */
int fsr_wgrad_c3(const float* img, const void* act, float* out, int N, int H, int W, int C64, int flip, int layout, int dtype,
                 void* stream);   /* layout 0: [27][C64]; 1: OIHW [3][C64][3][3] (head); 2: OIHW [C64][3][3][3] (necks) */
int fsr_bias_grad(const void* g, float* db, size_t npix, int C, int ps_perm, int dtype, void* stream);
int fsr_bias_grad_nchw(const float* g, float* db, int N, int C, size_t HW, void* stream);
/* torch.optim.AdamW (trainer.py:33-38,181,196) on flat fp32 buffers; g is multiplied by grad_scale first. */
int fsr_adamw(float* p, const float* g, float* m, float* v, size_t n, float lr, float b1, float b2, float eps, float wd,
              int step, float grad_scale, void* stream);

/* same, with the step counter in device memory (incremented by the call): nothing step-dependent is baked into the
 * launch parameters, so the whole train step can be captured once in a CUDA graph and replayed. */
int fsr_adamw_dev(float* p, const float* g, float* m, float* v, size_t n, float lr, float b1, float b2, float eps, float wd,
                  int* step_dev, float grad_scale, void* stream);

/* ---- whole Generator.forward (model.py:112-117) as one call: neck -> n_layers residual blocks ->
 * bottleneck + long skip -> 2 x (conv + pixel-shuffle + PReLU) -> head + tanh. */
typedef struct FsrGeneratorParams {
  int n_filters; /* 64 in this build */
  int n_layers;  /* <= FSR_MAX_LAYERS */
  int dtype;     /* FSR_F16 | FSR_BF16 */
  int reserved;
  const float* neck_w;     /* [64,3,3,3] fp32 OIHW      model.py:76 */
  const float* neck_b;     /* [64]                                   */
  const float* neck_alpha; /* [1]                       model.py:77 */
  const void* stem_w1[FSR_MAX_LAYERS];     /* packed [9][64][64]   model.py:47 */
  const float* stem_alpha[FSR_MAX_LAYERS]; /* [1]                  model.py:56 */
  const void* stem_w2[FSR_MAX_LAYERS];     /* packed               model.py:57 */
  const void* bott_w;                      /* packed               model.py:87 */
  const void* up_w[2];      /* packed ps_perm [9][256][64]          model.py:30 */
  const float* up_b[2];     /* permuted [256]                                  */
  const float* up_alpha[2]; /* [1]                                  model.py:37 */
  const void* head_w;       /* packed [9][16][64] (3 real rows)     model.py:103 */
  const float* head_b;      /* padded [16]                                      */
} FsrGeneratorParams;

size_t fsr_generator_workspace_bytes(int N, int H, int W, int n_filters, int n_layers);
/* x: fp32 NCHW [N,3,H,W] or uint8 NHWC [N,H,W,3]; y: fp32 NCHW [N,3,4H,4W] or uint8 NHWC [N,4H,4W,3].
 * group > 0 runs neck..bottleneck over `group` images at a time so the residual-chain working set
 * stays inside the 126 MB L2 (0 = whole batch at once). */
int fsr_generator_forward(const FsrGeneratorParams* prm, const void* x, void* y, void* workspace, size_t ws_bytes,
                          int N, int H, int W, int in_u8, int out_u8, int group, void* stream);

/* Generator.forward can split the batch into `parts` sub-batches (default 1, env FSR_STREAMS) on internal side streams
 * (fork/join by events on the caller's stream): the HBM-bound kernels of one sub-batch overlap the tensor-bound
 * convolutions of the other.  1 = single stream. */
int fsr_set_overlap_streams(int parts);

/* ---- measurement hooks (bench.py): device-time single kernels INSIDE a running forward.
 * fsr_profile_enable(kernel_id) makes every later launch of that kernel (FSR_K_*) be bracketed by a
 * cudaEvent pair on its launch stream (at most FSR_PROFILE_MAX pairs are kept);
 * fsr_profile_read() synchronises those events, writes the elapsed milliseconds and clears the list.
 * fsr_launch_count() = number of kernels this library has launched so far in this process. */
enum { FSR_K_NONE = -1, FSR_K_NECK = 0, FSR_K_CONV_RES = 1, FSR_K_IN_APPLY = 2, FSR_K_CONV_UP = 3,
       FSR_K_CONV_HEAD = 4, FSR_K_CONV_BIAS_ACT = 5, FSR_K_CONV_GEN = 6, FSR_K_CONV_WGRAD = 7 };
#define FSR_PROFILE_MAX 4096
int fsr_profile_enable(int kernel_id);
int fsr_profile_read(float* ms_out, int capacity);
/* same for several kernels at once: bit k of `mask` enables kernel id k; fsr_profile_read_ids also returns each record's id */
int fsr_profile_enable_mask(unsigned mask);
int fsr_profile_read_ids(float* ms_out, int* ids_out, int capacity);
/* ... and the ALGORITHMIC FLOPs of each timed launch (2*N*Ho*Wo*Cout*Cin*taps for the conv kernels, 0 otherwise) */
int fsr_profile_read_ex(float* ms_out, int* ids_out, double* flops_out, int capacity);
unsigned long long fsr_launch_count(void);

/* A-operand staging of the tensor-core conv: 0 = three column-shifted halo tiles per 8x16-pixel tile,
 * 1 = ONE halo tile per 16x8-pixel tile addressed through unaligned UMMA descriptors (2.6x less
 * L2->SM traffic, the default).  Environment FSR_HALO1=0 selects mode 0 at start-up. */
int fsr_set_halo_mode(int single_halo_tile);
/* 1: the 64-channel conv issues tcgen05.mma.ws (weight-stationary) pairs: the weight tile is read from smem once
 * per two pixel tiles (collector buffer reuse).  Default on (environment FSR_WS=0 disables; -1 = environment default). */
int fsr_set_ws_mode(int weight_stationary);

/* ---- rows SURVEY.md section 8 marks "next" --------------------------------------------------------------------
 * Validation metrics.  Replaces the torchmetrics objects of trainer.py:46-51 as used at trainer.py:60-68
 * (PeakSignalNoiseRatio / StructuralSimilarityIndexMeasure, data_range given, 11x11 gaussian sigma 1.5, reduction
 * "none"): ONE fused pass over pred / target (fp32 NCHW, device), each mapped v -> scale*v + shift first
 * (trainer.py:64-66: scale = shift = 0.5).  Accumulates (+=, double, device): sse[0] = sum (p-t)^2 over everything,
 * ssim_sum[n] = sum of image n's SSIM map over C x (H-10) x (W-10) valid window positions.  The caller forms
 * PSNR = 10 log10(dr^2 * numel / sse) and SSIM_n = ssim_sum[n] / (C (H-10)(W-10)).  taps11_host: 11 normalised gaussian
 * taps in HOST memory (copied into the launch parameters).  H, W >= 11, N*C <= 65535. */
int fsr_psnr_ssim(const float* pred, const float* target, int N, int C, int H, int W, float scale, float shift,
                  float data_range, const float* taps11_host, double* sse, double* ssim_sum, void* stream);

/* Data path.  Replaces NumpyImagesDataset.__getitem__ (dataloader.py:24-38) for a whole batch on a device-resident
 * uint8 image cache: sample b = (image index, crop_y, crop_x) -> hr[b] = crop/127.5 - 1 (fp32 NCHW [B,3,hr,hr],
 * hr = lr_size*scale) and lr[b] = antialiased-bicubic(crop -> lr_size x lr_size)/127.5 - 1 (torch
 * `_upsample_bicubic2d_aa`, what v2.Resize(antialias=True, BICUBIC) runs on a float tensor).  cache: uint8 CHW images back
 * to back, img_off/img_h/img_w per image; tap_min/tap_size/tap_w[lr_size][K]: the per-output-index tap table of the
 * hr -> lr resize (device memory; same table for rows and columns).  Out-of-range crops are clamped into the image. */
int fsr_crop_resize_aa(const uint8_t* cache, const int64_t* img_off, const int32_t* img_h, const int32_t* img_w,
                       const int32_t* samples, int B, int lr_size, int scale, const int32_t* tap_min,
                       const int32_t* tap_size, const float* tap_w, int K, float* lr, float* hr, void* stream);

/* The second conv of a ResidualBlock with the block's first InstanceNorm + PReLU fused into its load path:
 *   out, stats = conv3x3(PReLU(InstanceNorm(x_raw)))       (model.py:55-56 folded into :57-64)
 * x_raw [N,H,W,64] = RAW output of conv1 with its fixed-point statistics in_stats [N,64,2] (as written by
 * fsr_conv3x3_c64 RAW_STATS), in_alpha = PReLU slope (device), in_eps = InstanceNorm eps.  The normalised activation is
 * never written to HBM: the staged halo tile is transformed in shared memory (same fp32 operations as
 * fsr_instnorm_apply -> bit-identical results).  out must not alias x_raw.  Single-halo-tile mode only. */
int fsr_conv3x3_c64_in(const void* x_raw, const int64_t* in_stats, const float* in_alpha, float in_eps, const void* w_packed,
                       void* out, int64_t* stats, int N, int H, int W, int dtype, void* stream);
/* The FIRST conv of a ResidualBlock (or the bottleneck conv) with the PREVIOUS block's second InstanceNorm + skip fused
 * into its load path:
 *   x_next = InstanceNorm(x_raw) + res        (model.py:65 + :69 of block l; written to x_out for the next skip)
 *   out, stats = conv3x3(x_next)              (model.py:47-54 of block l+1, or the bottleneck :87-93)
 * x_raw = RAW conv2 output of block l with statistics in_stats; res = x_l; x_out = x_{l+1}; all [N,H,W,64] NHWC dtype.
 * x_out must alias neither res nor x_raw; out must alias none of them.  x_next, out and stats are bit-identical to
 * fsr_instnorm_apply(+residual) followed by fsr_conv3x3_c64 RAW_STATS.  Single-halo-tile mode only. */
int fsr_conv3x3_c64_res_in(const void* x_raw, const int64_t* in_stats, float in_eps, const void* res, void* x_out,
                           const void* w_packed, void* out, int64_t* stats, int N, int H, int W, int dtype, void* stream);
/* 1 (default): fsr_generator_forward runs the residual chain fully fused (fsr_conv3x3_c64_in + fsr_conv3x3_c64_res_in:
 * two launches per block, one InstanceNorm pass left per forward); 0: one normalise pass per block; -1: environment
 * default (FSR_FUSE_RES=0 disables).  Needs fuse_in. */
int fsr_set_fuse_res(int on);
/* 1 (default): the 64 -> 256 upsampling conv (FSR_EPI_PS_PRELU) runs as a CTA-pair kernel (tcgen05 cta_group::2, M = 256,
 * N = 256: each SM feeds the pair's MMA with its own pixel tile and HALF of the weights); 0: one CTA per 128-column half;
 * -1: environment default (FSR_UP_2CTA).  Bit-identical results. */
int fsr_set_up_2cta(int on);
/* 1 (default): fsr_generator_forward uses fsr_conv3x3_c64_in for every residual block; 0: separate normalise pass;
 * -1: environment default (FSR_FUSE_IN=0 disables). */
int fsr_set_fuse_in(int on);

/* 1 (default): fsr_conv3x3_gen runs weight-stationary over groups of four 128-pixel tiles (one weight fill per K step
 * and group instead of per tile: 2.3x less L2->smem traffic, same results bit for bit); 0: per-tile weight streaming;
 * -1: environment default (FSR_GEN_WS). */
int fsr_set_gen_ws(int on);
/* The same conv (stride 1, bias + activation epilogue, Cout % 128 == 0, W <= 25) on zero-bordered PADDED tensors
 * [N][H+2][W+2][C]: an M tile is 128 CONSECUTIVE positions of the flattened (n, y', x') index, so the <= 12x12 layers of
 * VGG19 (conv4_x, conv5_x behind model.py:8) fill 73 % / 56 % of every 128-row tile instead of 56 % / 28 % with 16x8
 * tiles.  Border positions of `out_padded` are written as zeros (the layout stays valid for the next layer).
 * mode 0 forward, mode 1 data gradient (weights packed with fsr_pack_conv3x3_weight_t).  fsr_maxpool2_padded /
 * fsr_maxpool2_relu_bwd_padded convert between the plain and the padded layout on the way (in_pad / out_pad = 0 | 1;
 * a padded output must be zero-filled by the caller: only interior pixels are written). */
int fsr_conv3x3_gen_flat(const void* x_padded, const void* w_packed, void* out_padded, const float* bias, int N, int H, int W,
                         int cin, int cout, int mode, int act, float slope, int dtype, void* stream);
int fsr_maxpool2_padded(const void* in, void* out, int N, int H, int W, int C, int in_pad, int out_pad, int dtype, void* stream);
int fsr_maxpool2_relu_bwd_padded(const void* in, const void* dout, void* din, int N, int H, int W, int C, int in_pad, int out_pad,
                                 int dtype, void* stream);
/* 1 (default): fsr_conv3x3_gen with Cout % 128 == 0 runs as a CTA-pair kernel (tcgen05 cta_group::2, M = 256, 128-wide
 * output-channel slices: half the shared-memory operand traffic per MMA); 0: the 64-wide single-CTA kernels;
 * -1: environment default (FSR_GEN_2CTA).  Same results bit for bit. */
int fsr_set_gen_2cta(int on);

/* 1 (default): the 3-channel-sided convs (fsr_neck_conv3x3, fsr_wgrad_c3) run on warp-level tensor-core MMAs
 * (mma.sync m16n8k16, fp32 operand split hi+lo: fp32-input accuracy); 0: the CUDA-core kernels (A/B and tests);
 * -1: environment default (FSR_SMALL_MMA). */
int fsr_set_small_mma(int on);

/* ---- contexts: caller-owned option overrides + internal side streams (re-entrancy across host threads / streams).
 * The fsr_set_* switches above are PROCESS-wide defaults.  A context carries its own values of the same switches and
 * its own side streams / events for the sub-batch overlap of fsr_generator_forward; fsr_ctx_bind makes it the calling
 * THREAD's current context for every later fsr_* call of that thread (NULL unbinds).  Two host threads, each with its
 * own context, stream, workspace and buffers, may drive the library concurrently (tests/test_capi_ctx_gpu.py).
 * fsr_ctx_set(ctx, option, value): value -1 = inherit the process default. */
enum { FSR_OPT_HALO1 = 0, FSR_OPT_WS = 1, FSR_OPT_FUSE_IN = 2, FSR_OPT_FUSE_RES = 3, FSR_OPT_UP_2CTA = 4, FSR_OPT_GEN_WS = 5,
       FSR_OPT_GEN_2CTA = 6, FSR_OPT_SMALL_MMA = 7, FSR_OPT_IN_BWD_FUSED = 8, FSR_OPT_OVERLAP_STREAMS = 9 };
int fsr_ctx_create(void** ctx_out);
int fsr_ctx_destroy(void* ctx);
int fsr_ctx_set(void* ctx, int option, int value);
int fsr_ctx_bind(void* ctx);

/* ---- PRECISE generator forward (Generator(compute_dtype=torch.float32); model.py:112-117 at ~fp32 accuracy).
 * The shipped checkpoint needs more than 16-bit operands for north_star's 1e-3 (fp16 measures 3.4e-3):
 * activations are stored fp32 NHWC and split a = a_hi + a_lo into two fp16 planes (fsr_split_f32 or the hi/lo outputs
 * below), weights likewise on the host, and every conv is THREE fsr_conv3x3_c64(..., FSR_EPI_F32) launches
 * (a_hi*w_hi store; a_lo*w_hi, a_hi*w_lo accumulate) into one fp32 NHWC buffer; the 64->3 head is three
 * fsr_conv3x3_c64(..., FSR_EPI_HEAD_TANH, out_u8 = 2 | 3 | 3) launches followed by fsr_tanh_f32. */
int fsr_split_f32(const float* x, void* hi, void* lo, size_t n_elems, void* stream);           /* n_elems % 8 == 0 */
/* neck (model.py:75-78) in fp32 arithmetic: x fp32 NCHW [N,3,H,W] -> out fp32 NHWC [N,H,W,64] = PReLU(conv + bias) */
int fsr_neck_conv3x3_f32(const float* x, const float* w, const float* bias, const float* alpha, float* out, int N, int H, int W,
                         void* stream);
/* InstanceNorm statistics (model.py:55,65,94) of an fp32 NHWC tensor [N,HW,64]: stats [N,64,2] int64 fixed point +=, caller zeroes */
int fsr_in_stats_f32(const float* x, int64_t* stats, int N, int HW, void* stream);

/* ---- n_filters = 32 generators (reference configs/config.yaml allows any n_filters; model.py:72-99) ----------------
 * A 32-channel NHWC tensor [N,H,W,32] (W even) is byte-identical to a 64-channel one on the "pixel-pair grid"
 * [N,H,W/2,64]; with the conv weights expanded on the host (fast_srgan_b200/pairs.py) the 64-channel kernels above
 * compute the 32-channel convolutions on it at half the zero-padding cost.  Two things differ on the pair grid:
 *  - fsr_in_stats_fold_pair: merge the InstanceNorm sums of slots c and 32+c (same channel, two pixel parities) of a
 *    stats block [N][64][2] written by FSR_EPI_RAW_STATS / fsr_conv3x3_c64_in, in place, so that the consumers
 *    (fsr_instnorm_apply, fsr_conv3x3_c64_in) normalise per real channel (model.py:14-24 InstanceNorm2d);
 *  - fsr_conv3x3_c64_head_pair: the 32->3 head + tanh (model.py:102-110) over pair rows; w_packed = pack of the
 *    expanded [6 (parity, rgb) -> 16][64][3][3] weight, bias[6]; out = fp32 NCHW [N,3,H,2*Wp] (out_u8 = 0) or uint8
 *    NHWC [N,H,2*Wp,3] (out_u8 = 1, inference.py:54-56). */
/* 3->32 neck conv + activation (model.py:75-78) storing 32-channel pixels [N,H,W,32] (= pair rows [N,H,W/2,64]);
 * w64/bias64 = the [32,3,3,3] weight and [32] bias zero-padded to 64 output channels. */
int fsr_neck_conv3x3_c32(const void* x, const float* w64, const float* bias64, const float* alpha, void* out, int N, int H,
                         int W, int act, float slope, int in_u8, int dtype, void* stream);
/* Performance hint for the CALLING THREAD's following fsr_conv3x3_c64 / _c64_in / _c64_head_pair launches: the packed
 * weights are pair-expanded, i.e. tap column 0 only has input slots 32..63 and tap column 2 only 0..31 non-zero, so 12 of
 * the 36 K-steps per tile are not issued.  Results are identical to on = 0 for such weights; switch it off afterwards. */
int fsr_set_pair_rows(int on);

/* Programmatic dependent launch for every kernel of the library (default OFF; env FSR_PDL=1): launches carry the
 * programmatic-stream-serialization attribute and every kernel waits for its predecessor's completion on the device
 * (griddepcontrol.wait) before its first global access - same results as plain stream order.  Measured slower on B200
 * for this workload, kept as a switch. */
int fsr_set_pdl(int on);
int fsr_in_stats_fold_pair(int64_t* stats, int N, void* stream);
int fsr_conv3x3_c64_head_pair(const void* x, const void* w_packed, void* out, const float* bias, int N, int H, int Wp,
                              int out_u8, int dtype, void* stream);
/* out = act((x - mean) * rstd) (+ residual), fp32 NHWC [N,HW,64]; hi/lo (both or neither): fp16 split planes of out */
int fsr_in_apply_f32(const float* x, const int64_t* stats, const float* residual, float* out, void* hi, void* lo,
                     const float* alpha, int act, int N, int HW, float eps, void* stream);
/* UpSamplingBlock tail (model.py:39-40): conv fp32 [N,H,W,256] (ps_perm column order) + bias_packed -> PixelShuffle(2) -> PReLU
 * -> out fp32 [N,2H,2W,64] (+ hi/lo planes) */
int fsr_ps_prelu_f32(const float* conv, const float* bias_packed, const float* alpha, float* out, void* hi, void* lo, int N, int H,
                     int W, void* stream);
/* tanh of the head's pre-activation fp32 NCHW [N,3,HW]: in place (out_u8 = NULL, model.py:109) or -> uint8 NHWC (inference.py:54-56) */
int fsr_tanh_f32(float* pre, uint8_t* out_u8, int N, int HW, void* stream);

/* ---- data-parallel exchange of the GAN step (the reference has no collective: trainer.py:180-181 and
 * :195-196 run on one device).  One process per GPU; the flat fp32 gradient buffer of a network is summed over ranks
 * with ncclAllReduce over NVLink/NVSwitch, issued on `stream` (capturable into the step's CUDA graph).  NCCL is bound
 * at run time (dlopen libnccl.so.2 - PyTorch's bundled copy when PyTorch is in the process); without it these return
 * FSR_ERR_NO_NCCL (-6).  Bootstrap: rank 0 calls fsr_nccl_unique_id, the 128 bytes travel to the other ranks by any
 * side channel (torch.distributed store / broadcast), every rank calls fsr_nccl_init with its CUDA device current. */
#define FSR_NCCL_ID_BYTES 128
int fsr_nccl_available(void);                                /* 1 if libnccl could be bound, else 0 */
int fsr_nccl_version(void);                                  /* e.g. 22809; <0 on error */
int fsr_nccl_unique_id(void* id_out_host);                   /* HOST buffer of FSR_NCCL_ID_BYTES */
int fsr_nccl_init(const void* id_host, int rank, int world, void** comm_out);     /* collective: all ranks call it */
/* in-place sum over ranks of n fp32 values at `buf` (device) - after trainer.py:180 (discriminator) / :195 (generator) */
int fsr_nccl_allreduce(void* comm, float* buf, size_t n, void* stream);
/* rank `root`'s buffer to every rank (replica initialisation / resume) */
int fsr_nccl_broadcast(void* comm, float* buf, size_t n, int root, void* stream);
int fsr_nccl_destroy(void* comm);

#ifdef __cplusplus
}
#endif
#endif /* FSR_B200_H_ */
