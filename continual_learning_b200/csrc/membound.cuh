// membound.cuh — host launchers of the HBM-bound kernels (membound.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace clk {

struct AdamTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long numel;
};

void set_num_sms(int n);

cudaError_t nchw_f32_to_nhwc_bf16(const float* x, void* y, int N, int C, int H, int W, int Cpad,
                                  cudaStream_t st);
cudaError_t nhwc_to_nchw_f32(const void* x, int x_is_f32, float* y, int N, int C, int H, int W, int ldc,
                             cudaStream_t st);
cudaError_t im2col3x3_stem(const float* x, void* a, int N, int Cin, int H, int W, cudaStream_t st);
cudaError_t pack_w(const float* src, void* outAB, void* outBA, int A, int B, int T, int ldA, int ldB,
                   int ldB2, int ldA2, int rev, cudaStream_t st);
cudaError_t unpack_wgrad(const float* D, float* grad, int A, int B, int T, int ldA, int ldB, float alpha,
                         int accumulate, int transposed, cudaStream_t st);
cudaError_t bn_finalize(const double* sum, const double* sq, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, float* mean_out, float* invstd_out,
                        float* scale, float* shift, int C, double count, float eps, float momentum,
                        int training, cudaStream_t st);
cudaError_t bn_apply(const void* y, void* z, const float* scale, const float* shift, long long P, int C,
                     cudaStream_t st);
cudaError_t bn_apply_pool(const void* y, void* z, void* pooled, void* idx, const float* scale,
                          const float* shift, int N, int H, int W, int C, cudaStream_t st);
cudaError_t maxpool_bwd_add_reduce(const void* dpooled, const void* idx, const void* skip, const void* y, void* din,
                                   double* s1, double* s2, int N, int H, int W, int C, cudaStream_t st);
cudaError_t maxpool_bwd_add(const void* dpooled, const void* idx, const void* skip, void* din, int N,
                            int H, int W, int C, cudaStream_t st);
cudaError_t bn_stats(const void* y, double* sum, double* sq, long long P, int C, cudaStream_t st);
cudaError_t bn_bwd_reduce(const void* dz, const void* y, double* s1, double* s2, long long P, int C,
                          cudaStream_t st);
cudaError_t bn_bwd_finalize(const double* s1, const double* s2, const float* gamma, const float* mean,
                            const float* invstd, float* dgamma, float* dbeta, float* kA, float* kB,
                            float* kC, int C, double count, int training, int accumulate,
                            cudaStream_t st);
cudaError_t bn_relu_bwd_apply(const void* dz, const void* y, void* dpre, const float* kA, const float* kB,
                              const float* kC, double* dbias, long long P, int C, cudaStream_t st);
cudaError_t channel_sum(const void* g, double* out, long long P, int C, cudaStream_t st);
cudaError_t f64_to_f32(const double* src, float* dst, int n, int ld_group, int groups, float alpha,
                       int accumulate, cudaStream_t st);
cudaError_t ce_kd_loss(const float* logits, const float* old_logits, const long long* labels, long long P,
                       int C, int Cold, float T, float lambda, float gscale, void* dlogits, int ldd,
                       double* loss_acc, int* err_flag, cudaStream_t st);
cudaError_t confusion_matrix(const long long* target, const long long* pred, long long n, int nc,
                             long long* conf, int* err_flag, cudaStream_t st);
cudaError_t argmax_confusion(const float* logits, const long long* labels, long long P, int C, int nc,
                             long long* pred_out, long long* conf, long long* correct, cudaStream_t st);
cudaError_t adam_multi_tensor(const AdamTensor* tensors, const void* blocks, int nblocks, int chunk,
                              float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt,
                              float gscale, const float* hyper, cudaStream_t st);
cudaError_t pack_w_multi(const void* jobs, int njobs, int total_tiles, int max_T, cudaStream_t st);
cudaError_t unpack_wgrad_multi(const void* jobs, int njobs, int total_tiles, int max_T, cudaStream_t st);
cudaError_t reduce_partials_multi(const void* jobs, int njobs, int total_blocks, cudaStream_t st);
cudaError_t f64_to_f32_multi(const void* jobs, int njobs, cudaStream_t st);

cudaError_t voc_prepare_batch(const void* items, int B, int H, int W, float* x, long long* y, int* err_flag,
                              cudaStream_t st);
cudaError_t labels_to_rgb(const long long* labels, long long n_images, long long hw, double* rgb, cudaStream_t st);

cudaError_t confusion_matrix_batched(const long long* target, const long long* pred, int B, long long n, int nc,
                                     long long* conf, int* err_flag, cudaStream_t st);

}  // namespace clk
