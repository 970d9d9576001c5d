/* clk.h — C ABI of libclk.so: the B200 (sm_100a) kernels behind the U-Net training step.
 *
 * The reference (LorenzoFramba/Continual-Learning) has no FFI of its own: its "operator API" for this
 * path is the stock PyTorch module calls made by models/unet.py and trainer.py.  Each entry point
 * below names the reference call site (file:line under /root/reference) whose arithmetic it replaces;
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns an int status (CLK_OK or a negative CLK_E_*), never throws;
 *     clk_last_error() returns a thread-local message for the last failure on this thread;
 *   - all tensor arguments are caller-owned DEVICE pointers; nothing persistent is allocated;
 *   - every launch is asynchronous on the given stream (a cudaStream_t passed as void*);
 *   - activations are NHWC bf16 ("pixels x channels", channel counts multiples of 64 unless stated),
 *     statistics / parameters / gradients are fp32, accumulators marked "f64" are double;
 *   - there is no fallback: a device that is not sm_100 makes clk_query_device() fail.
 */
#ifndef CLK_H_
#define CLK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLK_OK 0
#define CLK_E_BADARG (-1)
#define CLK_E_UNSUPPORTED_SHAPE (-2)
#define CLK_E_WORKSPACE (-3)
#define CLK_E_CUDA (-4)
#define CLK_E_ARCH (-5)

typedef void* clk_stream_t;

/* ---- library ---- */
int clk_version(void);
const char* clk_last_error(void);
/* CLK_OK iff device `dev` is compute capability 10.x; caches the SM count for grid sizing. */
int clk_query_device(int dev);
/* Developer knobs that select between kernel variants of the same operator (every variant computes the same
 * function and is covered by tests/test_gpu_kernels.py): "conv3_v2" (0 generic per-tap kernel, 2 one-CTA halo kernel,
 * 4 CTA-pair halo kernel = default), "conv3_pair", "conv3_rowtap" (1 = paired-tap kernel for 64-output-channel layers,
 * default), "fprop_bn" (force the N tile, 0 = auto), "wgrad_v2", "wgrad_bn", "wgrad_ksplit", "wgrad_ctas",
 * "conv3_min_hw", "convT_wide", "pdl", "pdl_tensor_trigger".
 * These are PROCESS-GLOBAL and not synchronised: they are the one exception to "no global state except the error
 * string" — set them before the first launch (or between launches of a single-threaded test), never concurrently
 * with launches from another thread.  Production code does not call this function. */
int clk_set_tuning(const char* key, int value);

/* ---- layout (trainer.py:168 inputs.to(device); models/unet.py:74 forward input/output) ---- */
int clk_nchw_f32_to_nhwc_bf16(const float* x, void* y, int N, int C, int H, int W, int Cpad,
                              clk_stream_t st);
int clk_nhwc_to_nchw_f32(const void* x, int x_is_f32, float* y, int N, int C, int H, int W, int ldc,
                         clk_stream_t st);
/* stem im2col for the Cin<=7 first conv (models/unet.py:50): A[N*H*W][64] bf16, k = c*9+r*3+s. */
int clk_im2col3x3_stem(const float* x_nchw, void* a, int N, int Cin, int H, int W, clk_stream_t st);
/* the whole stem layer enc1.0 (models/unet.py:50: Conv2d(3, 64, 3, padding=1) + ReLU, + BatchNorm statistics or the
 * inference affine) as ONE launch: the im2col tile is built in shared memory from the fp32 NCHW input, never written to
 * HBM (12 B in + 128 B out per pixel instead of the 134 MB matrix of clk_im2col3x3_stem + clk_gemm_fprop, which remain for
 * other shapes and for the stem's weight gradient).  x f32 [N][Cin][H][W], w = clk_pack_w'ed bf16 [64][64]
 * (k = c*9 + r*3 + s), y bf16 NHWC [N][H][W][64].  Needs Cin*9 <= 32, H % 8 == 0, W % 16 == 0. */
int clk_stem_conv3x3_fprop(const float* x, int Cin, const void* w, const float* bias, void* y, double* stat_sum,
                           double* stat_sq, const float* bn_scale, const float* bn_shift, int N, int H, int W, int relu,
                           clk_stream_t st);

/* ---- weights: fp32 PyTorch layout <-> packed bf16 operand layout ----
 * src fp32 [A][B][T] (Conv2d: A=Cout,B=Cin,T=kh*kw; ConvTranspose2d: A=Cin,B=Cout,T=4)
 * outAB bf16 [T][ldA][ldB], outBA bf16 [T][ldB2][ldA2] (tap order reversed when rev!=0); either may be NULL. */
int clk_pack_w(const float* src, void* outAB, void* outBA, int A, int B, int T, int ldA, int ldB,
               int ldB2, int ldA2, int rev, clk_stream_t st);
/* grad[A][B][T] (=|+=) alpha * D[T][ldA][ldB]   (D = packed fp32 weight gradient);
 * transposed != 0: D is [T][ldB][ldA] (the conv3x3 wgrad layout [9][Cin][Cout] with A = Cout, B = Cin). */
int clk_unpack_wgrad(const float* D, float* grad, int A, int B, int T, int ldA, int ldB, float alpha,
                     int accumulate, int transposed, clk_stream_t st);

/* Batched (table-driven) forms: ONE launch for every layer of the model. `jobs` is a device array of
 * int64[16] rows (pointers and ints widened to int64, alpha as the bit pattern of a double):
 *   pack:    {src, outAB, outBA, A, B, T, ldA, ldB, ldB2, ldA2, rev, tile0, tiles_b}
 *   unpack:  {D, grad, A, B, T, ldA, ldB, alpha, accumulate, tile0, tiles_b, transposed, nsplit, split_stride}
 *   convert: {src_f64, dst_f32, n, ld_group, groups, alpha, accumulate}   (one block per row)
 * tile0 = index of the job's first 32x32 tile in the launch grid, tiles_b = ceil(B/32). */
int clk_pack_w_multi(const void* jobs, int njobs, int total_tiles, int max_T, clk_stream_t st);
int clk_unpack_wgrad_multi(const void* jobs, int njobs, int total_tiles, int max_T, clk_stream_t st);
int clk_f64_to_f32_multi(const void* jobs, int njobs, clk_stream_t st);
/* split-K partial buffers -> their sum, in place in split 0, fixed order (deterministic).
 *   reduce:  {base, n_float4, nsplit, stride_float4, block0}  (256 float4 per block; block0 = first block of the job) */
int clk_reduce_partials_multi(const void* jobs, int njobs, int total_blocks, clk_stream_t st);

/* ---- tcgen05 implicit GEMMs ----
 * conv3x3, stride 1, zero pad 1 (nn.Conv2d at models/unet.py:13,16,28,31,53,66,69) fused with
 * bias + ReLU (models/unet.py:14,17,...) and the per-channel sum / sum-of-squares BatchNorm needs.
 * The input may be two tensors whose channels are concatenated (torch.cat at models/unet.py:83-87,
 * skip first): x0 has C0 channels, x1 has C1 (x1 may be NULL with C1 = 0).
 * w: bf16 [9][Cout][C0+C1] (clk_pack_w outAB).  y: bf16 [N][H][W][Cout].  stat_*: f64[Cout] or NULL. */
int clk_conv3x3_fprop(const void* x0, int C0, const void* x1, int C1, const void* w, const float* bias,
                      void* y, double* stat_sum, double* stat_sq, int N, int H, int W, int Cout, int relu,
                      clk_stream_t st);
/* Inference form (module.eval(), trainer.py:271, and the frozen old model of the continual step): BatchNorm with
 * running statistics is a per-channel affine map known before the conv runs, so it is applied in the epilogue:
 * z = relu(conv(x) + bias) * bn_scale + bn_shift, one launch, no intermediate tensor.  bn_scale / bn_shift: f32[Cout]
 * from clk_bn_finalize(training = 0). */
int clk_conv3x3_fprop_eval(const void* x0, int C0, const void* x1, int C1, const void* w, const float* bias,
                           void* z, const float* bn_scale, const float* bn_shift, int N, int H, int W, int Cout,
                           int relu, clk_stream_t st);
/* dgrad of the same conv (autograd of trainer.py:175): dy bf16 [N][H][W][Cout],
 * wd bf16 [9][C0+C1][Cout] (clk_pack_w outBA, rev=1); writes dx0 [..][C0] and dx1 [..][C1]. */
int clk_conv3x3_dgrad(const void* dy, int Cout, const void* wd, void* dx0, int C0, void* dx1, int C1,
                      int N, int H, int W, clk_stream_t st);
/* wgrad: dw fp32 [9][C0+C1][Cout] += sum_pixels shifted x (x) dy  (tap, input channel, output channel).
 * Caller zeroes dw; clk_unpack_wgrad(..., transposed=1) turns it into the PyTorch [Cout][Cin][3][3] layout. */
int clk_conv3x3_wgrad(const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1, float* dw,
                      int N, int H, int W, clk_stream_t st);
/* Deterministic split-K form used by the step: K split s stores its partial sums to partials + s*9*Cin*Cout with
 * plain stores (no atomics, nothing to zero); clk_conv3x3_wgrad_splits() returns the number of splits the kernel
 * will use for a shape (pure host function), and clk_unpack_wgrad_multi sums the partials in a fixed order. */
int clk_conv3x3_wgrad_splits(int Cout, int Cin, int N, int H, int W);
int clk_conv3x3_wgrad_split(const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1, float* partials,
                            int N, int H, int W, clk_stream_t st);
/* plain GEMM out[P][.] = a[P][K] * w[Npad][K]^T (+bias, ReLU, stats): the im2col'ed stem conv
 * (models/unet.py:50), the 1x1 head (models/unet.py:72; fp32 output, n_store = num_classes) and the
 * head dgrad.  K % 64 == 0; Npad in {32 (f32 out), 64, 128, 256 multiples}. */
int clk_gemm_fprop(const void* a, int K, const void* w, const float* bias, void* out, int ldo, int n_store,
                   int out_is_f32, int relu, double* stat_sum, double* stat_sq, long long P, int Npad,
                   clk_stream_t st);
/* inference form of the stem (bf16 output): out = relu(a * w^T + bias) * bn_scale + bn_shift */
int clk_gemm_fprop_eval(const void* a, int K, const void* w, const float* bias, void* out, int ldo, int n_store,
                        int relu, const float* bn_scale, const float* bn_shift, long long P, int Npad,
                        clk_stream_t st);
/* The 1x1 head (models/unet.py:72), nn.CrossEntropyLoss (trainer.py:113,174) [+ the distillation term of
 * clk_ce_kd_loss] and the head's share of loss.backward() (trainer.py:175) in ONE launch: the logits stay in tensor
 * memory, z is read once and dz written once.  z bf16 [P][64]; wf bf16 [32][64], wd bf16 [64][64] (clk_pack_w of the
 * head weight); bias f32[C] or NULL; labels int64[P]; old_logits f32 [P][Cold] or NULL.
 * dz bf16 [P][64] = gradient w.r.t. z; dw f32 [>=C][64] += dW; dbias f64[>=C] += db; loss_acc f64[2] += {sum CE, sum KL}
 * (same conventions as clk_ce_kd_loss; gscale folds the 1/P of the mean).  err_flag set on a label outside [0, C). */
int clk_head_loss_bwd(const void* z, const void* wf, const void* wd, const float* bias, const int64_t* labels,
                      const float* old_logits, long long P, int Cin, int C, int Cold, float T, float lambda,
                      float gscale, void* dz, float* dw, double* dbias, double* loss_acc, int* err_flag,
                      clk_stream_t st);
/* Statistics / validation path in ONE launch: the 1x1 head (models/unet.py:72), argmax over classes
 * (trainer.py:183,279), the correct-pixel count (trainer.py:184) and metrics._fast_conf_matrix (metrics.py:32-38)
 * without materialising the logits.  z bf16 [P][64]; wf bf16 [32][64]; bias f32[C] or NULL; labels int64[P];
 * pred_out int64[P] or NULL; conf int64[nc*nc] += (rows = target, targets outside [0, nc) skipped) or NULL;
 * correct int64[1] += or NULL.  Same results as clk_gemm_fprop (fp32 logits) + clk_argmax_confusion. */
int clk_head_argmax_confusion(const void* z, const void* wf, const float* bias, const int64_t* labels, long long P,
                              int Cin, int C, int nc, int64_t* pred_out, int64_t* conf, int64_t* correct,
                              clk_stream_t st);
/* out fp32 [ld_u][ld_t] += u[P][CU]^T * t[P][CT]  (weight gradient of the GEMMs above) */
int clk_gemm_wgrad(const void* u, int CU, const void* t, int CT, float* out, int ld_u, int ld_t,
                   long long P, clk_stream_t st);
/* ConvTranspose2d k=2 s=2 (models/unet.py:34): x bf16 [N][H][W][Cin], w bf16 [4*Cout][Cin]
 * (clk_pack_w outBA, rev=0), y bf16 [N][2H][2W][Cout] written through the pixel-shuffle epilogue. */
int clk_convT2x2_fprop(const void* x, const void* w, const float* bias, void* y, int N, int H, int W,
                       int Cin, int Cout, clk_stream_t st);
/* dx[N][H][W][Cin] from dy[N][2H][2W][Cout]; wd bf16 [4][Cin][Cout] (clk_pack_w outAB). */
int clk_convT2x2_dgrad(const void* dy, const void* wd, void* dx, int N, int H, int W, int Cin, int Cout,
                       clk_stream_t st);
/* dw fp32 [4][Cin][Cout] += ...; caller zeroes. */
int clk_convT2x2_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout,
                       clk_stream_t st);

/* ---- BatchNorm2d (models/unet.py:15,18,30,33,52,55,68,71), eps/momentum as nn defaults ---- */
int clk_bn_stats(const void* y, double* sum, double* sq, long long P, int C, clk_stream_t st);
/* training: batch mean / biased var -> mean, invstd, scale=gamma*invstd, shift=beta-mean*scale and the
 * running-stat EMA (unbiased var); eval (training=0): coefficients from the running stats. */
int clk_bn_finalize(const double* sum, const double* sq, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float* mean_out, float* invstd_out,
                    float* scale, float* shift, int C, double count, float eps, float momentum,
                    int training, clk_stream_t st);
int clk_bn_apply(const void* y, void* z, const float* scale, const float* shift, long long P, int C,
                 clk_stream_t st);
/* BN apply fused with MaxPool2d(2,2) (models/unet.py:12,80): z and pooled + 1-byte window index.
 * scale == NULL: pooling only (z untouched). */
int clk_bn_apply_pool(const void* y, void* z, void* pooled, void* idx, const float* scale,
                      const float* shift, int N, int H, int W, int C, clk_stream_t st);
/* Fused-finalize forms used by the step (one launch instead of two): the per-channel coefficients are computed
 * from the raw sums in the kernel prologue; block 0 writes mean / invstd / running stats (resp. dgamma / dbeta). */
/* din = scatter(dpooled by idx) + skip (skip may be NULL): max-pool backward fused with the
 * skip-connection gradient sum. */
int clk_maxpool_bwd_add(const void* dpooled, const void* idx, const void* skip, void* din, int N, int H,
                        int W, int C, clk_stream_t st);
/* the same, fused with the BatchNorm-backward reductions of the pooled layer (models/unet.py:12,15): din is that
 * layer's dz, so s1 f64[C] += sum din and s2 f64[C] += sum din * y (y = the layer's relu(conv) output, bf16 [N][H][W][C])
 * are accumulated on the values stored, and clk_bn_bwd_reduce need not run.  C / 8 must divide 256. */
int clk_maxpool_bwd_add_reduce(const void* dpooled, const void* idx, const void* skip, const void* y, void* din,
                               double* s1, double* s2, int N, int H, int W, int C, clk_stream_t st);
int clk_bn_bwd_reduce(const void* dz, const void* y, double* s1, double* s2, long long P, int C,
                      clk_stream_t st);
int clk_bn_bwd_finalize(const double* s1, const double* s2, const float* gamma, const float* mean,
                        const float* invstd, float* dgamma, float* dbeta, float* kA, float* kB, float* kC,
                        int C, double count, int training, int accumulate, clk_stream_t st);
/* dpre = relu'(y) * (kA*dz + kB*y + kC); dbias f64[C] += sum dpre */
int clk_bn_relu_bwd_apply(const void* dz, const void* y, void* dpre, const float* kA, const float* kB,
                          const float* kC, double* dbias, long long P, int C, clk_stream_t st);
int clk_channel_sum(const void* g, double* out, long long P, int C, clk_stream_t st);
int clk_f64_to_f32(const double* src, float* dst, int n, int ld_group, int groups, float alpha,
                   int accumulate, clk_stream_t st);

/* ---- loss: nn.CrossEntropyLoss (trainer.py:113,174) fused with the temperature-KL distillation
 * term of the continual step (not in the reference) and with its own backward ----
 * logits fp32 [P][C]; old_logits fp32 [P][Cold] or NULL; labels int64 [P];
 * dlogits bf16 [P][ldd] (columns >= C zero); loss_acc f64[2] += {sum CE, sum KD}. */
int clk_ce_kd_loss(const float* logits, const float* old_logits, const int64_t* labels, long long P, int C,
                   int Cold, float T, float lambda, float gscale, void* dlogits, int ldd, double* loss_acc,
                   int* err_flag, clk_stream_t st);

/* ---- metrics (metrics.py:32-38,55-63; trainer.py:183-184,279-280) ---- */
/* conf int64 [nc*nc] += bincount(nc*t+p) over 0<=t<nc; *err_flag=1 if a kept target has p outside [0,nc). */
int clk_confusion_matrix(const int64_t* target, const int64_t* pred, long long n, int nc, int64_t* conf,
                         int* err_flag, clk_stream_t st);
/* argmax over fp32 logits [P][C] + confusion + correct count in one pass; any output may be NULL. */
int clk_argmax_confusion(const float* logits, const int64_t* labels, long long P, int C, int nc,
                         int64_t* pred_out, int64_t* conf, int64_t* correct, clk_stream_t st);

/* ---- optimiser: torch.optim.Adam step (trainer.py:108-110,176) over a table of tensors ----
 * tensors: device array of {float* p; const float* g; float* m; float* v; int64 numel};
 * blocks: device array of int2 {tensor index, chunk index}; bc1 = 1-b1^t, bc2_sqrt = sqrt(1-b2^t).
 * hyper_dev (optional): device float[4] {lr, bc1, bc2_sqrt, gscale} overriding the by-value scalars, so a
 * captured CUDA graph of the step can be replayed with a new step count / learning rate. */
int clk_adam_multi_tensor(const void* tensors, const void* blocks, int nblocks, int chunk, float lr,
                          float b1, float b2, float eps, float bc1, float bc2_sqrt, float gscale,
                          const float* hyper_dev, clk_stream_t st);

/* ---- data contract on either side of the step ----
 * One VOC.__getitem__ per sample after image decoding (datasets/voc.py:127-140 with the transform of main.py:17-23):
 * Pad(10) + CenterCrop((H, W)) expressed as the crop origin (top, left) in source coordinates (zero fill outside the
 * source; the host computes it with torchvision's rounding), ToTensor + Normalize(0.5, 0.5) on the image and to_mask
 * (datasets/voc.py:56-72: palette RGB -> class index, void -> 0) on the mask.
 * items: device int64[B][8] rows {img u8 [Hs][Ws][3], mask u8 [Hs][Ws][3], Hs, Ws, top, left, 0, 0} (either pointer
 * may be 0).  x: f32 [B][3][H][W] bit-equal to the reference's; y: int64 [B][H][W]; a colour that is not in the
 * palette (the reference raises ValueError) writes -1 and sets *err_flag. */
int clk_voc_prepare_batch(const void* items, int B, int H, int W, float* x, int64_t* y, int* err_flag,
                          clk_stream_t st);
/* datasets/voc.py:74-89 (to_rgb, used by the sample dump trainer.py:193-194): labels int64 [B][hw] ->
 * f64 [B][3][hw] palette colours; indices outside [0, 22) keep their value in all three channels. */
int clk_labels_to_rgb(const int64_t* labels, long long n_images, long long hw, double* rgb, clk_stream_t st);

/* One confusion matrix PER IMAGE: the counts behind the legacy per-image metrics
 * pixel_accuracy / mean_accuracy / mean_IU / frequency_weighted_IU (metrics.py:74-183).  target, pred: int64 [B][n];
 * conf: int64 [B][nc*nc] += (rows = target); a value outside [0, nc) in either map sets *err_flag. */
int clk_confusion_matrix_batched(const int64_t* target, const int64_t* pred, int B, long long n, int nc, int64_t* conf,
                                 int* err_flag, clk_stream_t st);

#ifdef __cplusplus
}
#endif
#endif /* CLK_H_ */
