// igemm.cuh — parameter blocks and host launchers of the tcgen05 implicit-GEMM kernels.
//
// Two kernels cover every dense contraction of the U-Net step:
//   * igemm_fprop: D[pixels, n] = sum_{tap, k} A[pixel (+) tap, k] * B[tap][n][k]
//       A = NHWC bf16 activations (K-major via TMA, zero fill outside the image = conv padding),
//       up to two A sources whose channels are walked back to back (skip-connection concat folded
//       into the addressing, reference models/unet.py:83-87), B = packed bf16 weights.
//       Used for conv3x3 forward + dgrad, ConvTranspose2d forward + dgrad, conv1x1 forward + dgrad.
//   * igemm_wgrad: D[tap][u, t] = sum_{pixels} U[pixel, u] * T[pixel (+) tap, t]
//       both operands MN-major straight out of NHWC memory; split-K over pixel tiles with fp32
//       red.global accumulation.  Used for every weight gradient.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace clk {

enum AddrMode : int {
  ADDR_LINEAR = 0,  // A is [P][C]; tile = consecutive pixels; no taps
  ADDR_NHWC = 1,    // A is [N][H][W][C]; tile = nb x th x tw pixels; taps shift (h, w); OOB -> 0
  ADDR_QUAD = 2     // A is the stride-2 quadrant view [N*H][2][W][2][C] of a [N][2H][2W][C] tensor
};

constexpr int kMaxTaps = 9;

struct TileGeom {
  int mode;                       // AddrMode
  int tw, th, nb;                 // tile = nb*th*tw pixels (ADDR_QUAD: th rows of the merged N*H axis, nb=1)
  int tiles_w, tiles_h, tiles_n;  // tile grid (LINEAR: tiles_w only; QUAD: tiles_w x tiles_h)
  int N, H, W;                    // pixel grid of the tiled tensor
  int t1[kMaxTaps], t2[kMaxTaps], t3[kMaxTaps];  // per-tap offsets added to TMA coords 1..3
};

struct FpropParams {
  TileGeom g;
  int ntaps;
  int kc0, kc1;        // 64-channel K chunks taken from A source 0 / source 1
  // epilogue
  int shuffle;         // ConvTranspose2d pixel shuffle: column = quadrant*cout_q + o
  int cout_q;
  int n_store;         // columns >= n_store are not stored (padding columns)
  int split_c;         // columns >= split_c go to dst1 (0 = single destination)
  void* dst0;
  void* dst1;
  int ldc0, ldc1;      // row pitch (elements) of dst0 / dst1
  const float* bias;   // per output channel (per o when shuffle); may be null
  int relu;
  double* stat_sum;    // per-column sum / sum of squares of the stored values; may be null
  double* stat_sq;
  const float* bn_scale;  // inference: out = relu(acc + bias) * bn_scale + bn_shift (stat_* must then be null)
  const float* bn_shift;
  // bf16 output through shared-memory staging + TMA tensor stores (full 128-byte lines, clipped at the tensor border)
  // instead of per-thread 16-byte stores at a >= 128-byte stride.  0 = off; 1 = ADDR_LINEAR rows map {C, P};
  // 2 = ConvTranspose2d pixel shuffle through the quadrant map {C, 2, 2W.., 2, N*H} of the output (nb == 1, H % th == 0);
  // 3 = ADDR_QUAD geometry, plain rows map {C, W, N*H}.  Needs n_store % 64 == 0 and a single destination.
  int tma_store;
};

struct WgradParams {
  TileGeom g;          // tile = 64 pixels
  int ntaps;           // all taps of the operator
  int G;               // taps accumulated per CTA (<= 4, G*BN <= 512)
  int CU;              // valid channels of U (rows of D)
  int CT;              // channels of T over both sources (columns of D)
  int ct_split;        // channels that live in T source 0
  int m_tiles, n_tiles, tap_groups;
  int ksplit, tiles_total;
  float* out;          // [ntaps][ld_u][ld_t] fp32 (or [ntaps][ld_t][ld_u] when transpose_out), red.add accumulated
  int ld_u, ld_t;
  int transpose_out;
};

void igemm_set_num_sms(int n);
cudaError_t igemm_set_pdl_mode(int mode);

// Launchers (igemm.cu). Return cudaError_t from the launch; maps are built by the caller.
cudaError_t launch_fprop(int BN, int out_is_f32, const CUtensorMap& a0, const CUtensorMap& a1,
                         const CUtensorMap& b, const FpropParams& p, int m_tiles, int n_tiles,
                         cudaStream_t st, const CUtensorMap* o = nullptr);
cudaError_t launch_wgrad(int BN, const CUtensorMap& u, const CUtensorMap& t0, const CUtensorMap& t1,
                         const WgradParams& p, cudaStream_t st);

// conv3x3 weight gradient, halo variant: one CTA owns (64 input channels) x (64 output channels) x ALL
// nine taps; per 16x8-pixel K tile it loads the dY tile once and one 18x16-pixel halo tile of X, and
// reads the nine shifted windows of the halo through row-shifted UMMA descriptors (two taps stacked in
// the M=128 rows of each MMA).
struct Wgrad9Params {
  int N, H, W;
  int tiles_w, tiles_h, tiles_total;  // 16 (h) x 8 (w) pixel K tiles
  int cin_slabs, split_slabs;         // 64-channel slabs of X over both sources / in source 0
  int cout_tiles;                     // 64-channel tiles of dY
  int ksplit;
  int Cin, Cout;
  float* out;                         // [9][Cin][Cout] fp32: red.add accumulated (caller zeroes) when split_stride == 0,
                                      // else split s stores its partial sums at out + s*split_stride (no atomics)
  long long split_stride;
};
// CTA-pair variant (Cout % 128 == 0): p.cout_tiles counts 128-channel tiles
cudaError_t launch_wgrad9x2(const CUtensorMap& u, const CUtensorMap& t0, const CUtensorMap& t1,
                            const Wgrad9Params& p, cudaStream_t st);
cudaError_t launch_wgrad9(const CUtensorMap& u, const CUtensorMap& t0, const CUtensorMap& t1,
                          const Wgrad9Params& p, cudaStream_t st);

// conv3x3 forward / dgrad, halo variant: persistent CTAs, 16x16-pixel super tile (two M=128 sub tiles),
// ONE 18x24-pixel halo load per 64-channel chunk feeds all nine taps through row-shifted descriptors,
// weights stream through their own ring, two TMEM accumulator stages overlap epilogue and main loop.
struct Conv3Params {
  int N, H, W;
  int tiles_w, tiles_h, m_tiles, n_tiles;
  int kc0, kc1;        // 64-channel chunks of source 0 / source 1
  int n_store, split_c;
  void* dst0;
  void* dst1;
  int ldc0, ldc1;
  const float* bias;
  int relu;
  double* stat_sum;
  double* stat_sq;
  const float* bn_scale;  // inference epilogue, see FpropParams
  const float* bn_shift;
};
// CTA-pair variant (tcgen05.mma.cta_group::2, M = 256): a0/a1 box {64, 16, 18}, b box {64, BN/2, 1}
cudaError_t launch_conv3x2(int BN, int SUB, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                           const Conv3Params& p, int num_sms, cudaStream_t st);
// row-tap variant for 64 output channels (three horizontal taps per MMA, N = 192), see igemm_conv3r_kernel
cudaError_t launch_conv3r(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& o,
                          const Conv3Params& p, int num_sms, cudaStream_t st);
cudaError_t launch_conv3(int BN, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                         const Conv3Params& p, int num_sms, cudaStream_t st);

// direct stem conv (3-channel fp32 NCHW input -> 64 channels bf16 NHWC), see stem_conv_kernel
cudaError_t launch_stem_conv(const CUtensorMap& w, const CUtensorMap& o, const float* x, int N, int Cin, int H, int W,
                             const float* bias, int relu, double* stat_sum, double* stat_sq, const float* bn_scale,
                             const float* bn_shift, int num_sms, cudaStream_t st);

// fused 1x1 head + softmax cross-entropy (+ distillation) + head backward, see head_loss_kernel
struct HeadLossParams {
  long long P;
  int C, Cold;
  float T, lambda, gscale;
  const float* bias;
  const long long* labels;
  const float* old_logits;   // fp32 [P][Cold] or null
  void* dz;                  // bf16 [P][64]
  float* dw;                 // fp32 [>= C][64], accumulated
  double* dbias;             // f64[>= C], accumulated
  double* loss_acc;          // f64[2], accumulated
  int* err_flag;
};
// z box {64, 128}; wf box {64, 32, 1}; wd box {64, 64, 1}
cudaError_t launch_head_loss(const CUtensorMap& z, const CUtensorMap& wf, const CUtensorMap& wd, const HeadLossParams& p,
                             int num_sms, cudaStream_t st);

// fused 1x1 head + arg-max + confusion matrix (statistics / validation path), see head_argmax_kernel
struct HeadArgmaxParams {
  long long P;
  int C, nc;
  const float* bias;
  const long long* labels;
  long long* pred_out;             // [P] or null
  unsigned long long* conf;        // [nc * nc] accumulated, or null
  unsigned long long* correct;     // accumulated count of pred == label, or null
};
// z box {64, 128}; wf box {64, 32, 1}
cudaError_t launch_head_argmax(const CUtensorMap& z, const CUtensorMap& wf, const HeadArgmaxParams& p, int num_sms,
                               cudaStream_t st);

}  // namespace clk
