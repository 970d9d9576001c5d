// membound.cu — the HBM-bound kernels of the U-Net step (sm_100a): layout conversion, im2col of
// the 3-channel stem, weight (un)packing, BatchNorm statistics/apply/backward, max-pool with
// indices, the fused cross-entropy + distillation loss, argmax/confusion-matrix histogram and
// multi-tensor Adam.  All activations are NHWC bf16; every kernel moves 16-byte vectors, uses
// grid-stride loops over grids sized in multiples of the SM count, and reduces with warp shuffles /
// shared memory before touching global atomics.
//
// Reference ops replaced (file:line in /root/reference): nn.BatchNorm2d models/unet.py:15,18,30,33,
// 52,55,68,71; nn.MaxPool2d models/unet.py:12,80; nn.CrossEntropyLoss trainer.py:113,174;
// argmax/eq-count trainer.py:183-184; metrics._fast_conf_matrix metrics.py:32-38; optim.Adam
// trainer.py:108-110,176.
#include "membound.cuh"
#include "clk_ptx.cuh"

#include <cuda_bf16.h>
#include <math_constants.h>

namespace clk {

static int g_num_sms = 148;
void set_num_sms(int n) { g_num_sms = n > 0 ? n : 148; }
static inline int grid_for(long long work_items, int per_block, int waves = 8) {
  long long b = (work_items + per_block - 1) / per_block;
  long long cap = static_cast<long long>(g_num_sms) * waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf_lo(v.x); f[1] = bf_hi(v.x); f[2] = bf_lo(v.y); f[3] = bf_hi(v.y);
  f[4] = bf_lo(v.z); f[5] = bf_hi(v.z); f[6] = bf_lo(v.w); f[7] = bf_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ------------------------------------------------------------------------------------------
// layout conversion
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                             int N, int C, int HW, int Cpad) {
  pdl_launch_dependents();
  pdl_wait();
  // one thread per (n, pixel); channels are few in the use cases (stem input, logits)
  const long long total = static_cast<long long>(N) * HW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / HW, p = i % HW;
    for (int c = 0; c < Cpad; ++c) {
      float v = c < C ? x[(n * C + c) * HW + p] : 0.f;
      y[i * Cpad + c] = __float2bfloat16_rn(v);
    }
  }
}
__global__ void nhwc_bf16_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y,
                                             int N, int C, int HW, int ldc) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float tile[32][33];
  // grid: (ceil(HW/32), ceil(C/32), N); transposes a 32 pixel x 32 channel tile
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    tile[r][threadIdx.x] =
        (p < HW && c < C) ? __bfloat162float(x[(static_cast<long long>(n) * HW + p) * ldc + c]) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    if (p < HW && c < C) y[(static_cast<long long>(n) * C + c) * HW + p] = tile[threadIdx.x][r];
  }
}
__global__ void nhwc_f32_to_nchw_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int N,
                                            int C, int HW, int ldc) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (p < HW && c < C) ? x[(static_cast<long long>(n) * HW + p) * ldc + c] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    if (p < HW && c < C) y[(static_cast<long long>(n) * C + c) * HW + p] = tile[threadIdx.x][r];
  }
}

cudaError_t nchw_f32_to_nhwc_bf16(const float* x, void* y, int N, int C, int H, int W, int Cpad,
                                  cudaStream_t st) {
  const long long total = static_cast<long long>(N) * H * W;
  launch_k(nchw_f32_to_nhwc_bf16_kernel, dim3(grid_for(total, 256)), dim3(256), 0, st, 
      x, static_cast<__nv_bfloat16*>(y), N, C, H * W, Cpad);
  return cudaGetLastError();
}
cudaError_t nhwc_to_nchw_f32(const void* x, int x_is_f32, float* y, int N, int C, int H, int W,
                             int ldc, cudaStream_t st) {
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, N), block(32, 8);
  if (x_is_f32)
    launch_k(nhwc_f32_to_nchw_f32_kernel, dim3(grid), dim3(block), 0, st, static_cast<const float*>(x), y, N, C, H * W, ldc);
  else
    launch_k(nhwc_bf16_to_nchw_f32_kernel, dim3(grid), dim3(block), 0, st, static_cast<const __nv_bfloat16*>(x), y, N, C,
                                                         H * W, ldc);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// stem im2col: x NCHW fp32 [N,Cin,H,W] -> A [N*H*W][64] bf16, k = c*9 + r*3 + s (zero padded)
__global__ void im2col3x3_stem_kernel(const float* __restrict__ x, uint4* __restrict__ a, int N, int Cin,
                                      int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(N) * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(i % W);
    const int h = static_cast<int>((i / W) % H);
    const long long n = i / (static_cast<long long>(W) * H);
    float v[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) v[k] = 0.f;
    for (int c = 0; c < Cin; ++c) {
      const float* xc = x + (n * Cin + c) * H * W;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hh = h + r - 1;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int ww = w + s - 1;
          float t = 0.f;
          if (hh >= 0 && hh < H && ww >= 0 && ww < W) t = __ldg(xc + static_cast<long long>(hh) * W + ww);
          v[c * 9 + r * 3 + s] = t;  // runtime c: v lives in local memory (rare path)
        }
      }
    }
    uint4* o = a + i * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = v[j * 8 + e];
      o[j] = pack8(f);
    }
  }
}
// Cin == 3 specialisation keeps everything in registers (fully unrolled indices)
__global__ void im2col3x3_stem3_kernel(const float* __restrict__ x, uint4* __restrict__ a, int N, int H,
                                       int W) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(N) * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(i % W);
    const int h = static_cast<int>((i / W) % H);
    const long long n = i / (static_cast<long long>(W) * H);
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* xc = x + (n * 3 + c) * H * W;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hh = h + r - 1;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int ww = w + s - 1;
          float t = 0.f;
          if (hh >= 0 && hh < H && ww >= 0 && ww < W) t = __ldg(xc + static_cast<long long>(hh) * W + ww);
          v[c * 9 + r * 3 + s] = t;
        }
      }
    }
    uint4* o = a + i * 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = v[j * 8 + e];
      o[j] = pack8(f);
    }
#pragma unroll
    for (int j = 4; j < 8; ++j) o[j] = make_uint4(0, 0, 0, 0);
  }
}
cudaError_t im2col3x3_stem(const float* x, void* a, int N, int Cin, int H, int W, cudaStream_t st) {
  const long long total = static_cast<long long>(N) * H * W;
  if (Cin == 3)
    launch_k(im2col3x3_stem3_kernel, dim3(grid_for(total, 256, 16)), dim3(256), 0, st, x, static_cast<uint4*>(a), N, H, W);
  else
    launch_k(im2col3x3_stem_kernel, dim3(grid_for(total, 128, 16)), dim3(128), 0, st, x, static_cast<uint4*>(a), N, Cin, H, W);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// weight packing: src fp32 [A][B][T]  ->  outAB bf16 [T][ldA][ldB] (outAB[t][a][b] = src[a][b][t])
//                                     ->  outBA bf16 [T][ldB2][ldA2] (outBA[tt][b][a], tt = rev ? T-1-t : t)
__global__ void pack_w_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ outAB,
                              __nv_bfloat16* __restrict__ outBA, int A, int B, int T, int ldA, int ldB,
                              int ldB2, int ldA2, int rev) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float tile[];  // [T][32][33]
  const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  const int tid = threadIdx.x;  // 256 threads
  // coalesced read of rows a0..a0+31: each row has 32*T contiguous floats starting at b0*T
  const int rowlen = 32 * T;
  for (int idx = tid; idx < 32 * rowlen; idx += blockDim.x) {
    const int ar = idx / rowlen, rem = idx % rowlen;
    const int br = rem / T, t = rem % T;
    float v = 0.f;
    if (a0 + ar < A && b0 + br < B) v = src[(static_cast<long long>(a0 + ar) * B + b0 + br) * T + t];
    tile[(t * 32 + ar) * 33 + br] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < T * 1024; idx += blockDim.x) {
    const int t = idx / 1024, r = (idx / 32) % 32, c = idx % 32;
    if (outAB != nullptr && a0 + r < A && b0 + c < B)
      outAB[(static_cast<long long>(t) * ldA + a0 + r) * ldB + b0 + c] =
          __float2bfloat16_rn(tile[(t * 32 + r) * 33 + c]);
    if (outBA != nullptr && b0 + r < B && a0 + c < A) {
      const int tt = rev ? T - 1 - t : t;
      outBA[(static_cast<long long>(tt) * ldB2 + b0 + r) * ldA2 + a0 + c] =
          __float2bfloat16_rn(tile[(t * 32 + c) * 33 + r]);
    }
  }
}
cudaError_t pack_w(const float* src, void* outAB, void* outBA, int A, int B, int T, int ldA, int ldB,
                   int ldB2, int ldA2, int rev, cudaStream_t st) {
  dim3 grid((B + 31) / 32, (A + 31) / 32);
  const size_t smem = static_cast<size_t>(T) * 32 * 33 * sizeof(float);
  launch_k(pack_w_kernel, dim3(grid), dim3(256), smem, st, src, static_cast<__nv_bfloat16*>(outAB),
                                         static_cast<__nv_bfloat16*>(outBA), A, B, T, ldA, ldB, ldB2, ldA2,
                                         rev);
  return cudaGetLastError();
}

// packed fp32 gradient D[T][ldA][ldB] -> grad[A][B][T] (dst = alpha*D, or dst += alpha*D)
__global__ void unpack_wgrad_kernel(const float* __restrict__ D, float* __restrict__ grad, int A, int B,
                                    int T, int ldA, int ldB, float alpha, int accumulate, int transposed) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float tile[];  // [32 a][32*T + 1]
  const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  const int pitch = 32 * T + 1;
  for (int idx = threadIdx.x; idx < T * 1024; idx += blockDim.x) {
    const int t = idx / 1024, r = (idx / 32) % 32, c = idx % 32;
    if (transposed) {  // D is [T][ldB][ldA]: r walks b, c walks a
      float v = 0.f;
      if (a0 + c < A && b0 + r < B) v = D[(static_cast<long long>(t) * ldB + b0 + r) * ldA + a0 + c];
      tile[c * pitch + r * T + t] = v;
    } else {
      float v = 0.f;
      if (a0 + r < A && b0 + c < B) v = D[(static_cast<long long>(t) * ldA + a0 + r) * ldB + b0 + c];
      tile[r * pitch + c * T + t] = v;
    }
  }
  __syncthreads();
  const int rowlen = 32 * T;
  for (int idx = threadIdx.x; idx < 32 * rowlen; idx += blockDim.x) {
    const int ar = idx / rowlen, rem = idx % rowlen;
    const int br = rem / T;
    if (a0 + ar < A && b0 + br < B) {
      float* g = grad + (static_cast<long long>(a0 + ar) * B + b0) * T + rem;
      const float v = alpha * tile[ar * pitch + rem];
      *g = accumulate ? *g + v : v;
    }
  }
}
cudaError_t unpack_wgrad(const float* D, float* grad, int A, int B, int T, int ldA, int ldB, float alpha,
                         int accumulate, int transposed, cudaStream_t st) {
  dim3 grid((B + 31) / 32, (A + 31) / 32);
  const size_t smem = static_cast<size_t>(32) * (32 * T + 1) * sizeof(float);
  launch_k(unpack_wgrad_kernel, dim3(grid), dim3(256), smem, st, D, grad, A, B, T, ldA, ldB, alpha, accumulate, transposed);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// BatchNorm: statistics finalize (training) / eval coefficients
__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sq,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                   float* __restrict__ scale, float* __restrict__ shift, int C,
                                   double count, float eps, float momentum, int training) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, invstd;
  if (training) {
    const double m = sum[c] / count;
    double var = sq[c] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean = static_cast<float>(m);
    invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    if (running_mean != nullptr) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
    }
  } else {
    mean = running_mean[c];
    invstd = 1.f / sqrtf(running_var[c] + eps);
  }
  mean_out[c] = mean;
  invstd_out[c] = invstd;
  const float a = gamma[c] * invstd;
  scale[c] = a;
  shift[c] = beta[c] - mean * a;
}
cudaError_t bn_finalize(const double* sum, const double* sq, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, float* mean_out, float* invstd_out,
                        float* scale, float* shift, int C, double count, float eps, float momentum,
                        int training, cudaStream_t st) {
  launch_k(bn_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, st, sum, sq, gamma, beta, running_mean, running_var,
                                                      mean_out, invstd_out, scale, shift, C, count, eps,
                                                      momentum, training);
  return cudaGetLastError();
}

// z = scale[c]*y + shift[c]   (NHWC bf16, C % 8 == 0).  blockDim (256) is a multiple of CV, so the channel
// column of a thread never changes: coefficients live in registers; two 16-byte loads in flight per thread.
__global__ void __launch_bounds__(256) bn_apply_kernel(const uint4* __restrict__ y, uint4* __restrict__ z,
                                                        const float* __restrict__ scale,
                                                        const float* __restrict__ shift, long long nvec, int CV) {
  pdl_launch_dependents();
  pdl_wait();
  const int cv = CV - 1 - (threadIdx.x % CV);  // mirrored traversal (see below)
  float a[8], c[8];
  {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(scale) + cv * 2);
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(scale) + cv * 2 + 1);
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(shift) + cv * 2);
    const float4 c1 = __ldg(reinterpret_cast<const float4*>(shift) + cv * 2 + 1);
    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
    c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
  }
  // descending order: the conv epilogue wrote y in ascending tile order (its tail is still in L2), and the next
  // conv reads z in ascending order (the head, written last here, is then still in L2).  nvec is a multiple of
  // CV and so is the stride, hence the mirrored index keeps the thread's channel column (CV-1-cv) fixed.
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < nvec; j += 2 * stride) {
    const bool two = j + stride < nvec;
    const long long i = nvec - 1 - j - (two ? stride : 0);
    const uint4 v0 = ldg_stream(y + i);
    const uint4 v1 = two ? ldg_stream(y + i + stride) : make_uint4(0, 0, 0, 0);
    float f[8];
    unpack8(v0, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], a[e], c[e]);
    z[i] = pack8(f);
    if (two) {
      unpack8(v1, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], a[e], c[e]);
      z[i + stride] = pack8(f);
    }
  }
}
cudaError_t bn_apply(const void* y, void* z, const float* scale, const float* shift, long long P, int C,
                     cudaStream_t st) {
  const long long nvec = P * (C / 8);
  const int CV = C / 8;
  if (CV > 256 || CV < 1) return cudaErrorInvalidValue;
  const int threads = (256 / CV) * CV;  // a multiple of CV: the channel column of a thread is then constant
  launch_k(bn_apply_kernel, dim3(grid_for(nvec, threads * 4, 8)), dim3(threads), 0, st, static_cast<const uint4*>(y),
                                                                      static_cast<uint4*>(z), scale, shift, nvec, CV);
  return cudaGetLastError();
}

// z = scale*y + shift, pooled = maxpool2x2(z) with window index (0..3, first max in row-major
// window order, NaN propagates — the ATen rule), one thread per pooled pixel x 8 channels.
template <bool kApply>
__global__ void bn_apply_pool_kernel(const uint4* __restrict__ y, uint4* __restrict__ z,
                                     uint4* __restrict__ pooled, uint2* __restrict__ idx,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     int N, int H, int W, int CV) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(N) * Ho * Wo * CV;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % CV);
    long long r = i / CV;
    const int wo = static_cast<int>(r % Wo);
    r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    const long long n = r / Ho;
    float a[8], c[8];
    if (kApply) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(scale) + cv * 2);
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(scale) + cv * 2 + 1);
      const float4 c0 = __ldg(reinterpret_cast<const float4*>(shift) + cv * 2);
      const float4 c1 = __ldg(reinterpret_cast<const float4*>(shift) + cv * 2 + 1);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
    }
    float best[8];
    uint32_t bidx[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long pix = (n * H + 2 * ho + (k >> 1)) * W + 2 * wo + (k & 1);
      float f[8];
      unpack8(ldg_stream(y + pix * CV + cv), f);
      if (kApply) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], a[e], c[e]);
        const uint4 zz = pack8(f);
        z[pix * CV + cv] = zz;
        unpack8(zz, f);  // pool the values exactly as stored
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (k == 0 || f[e] > best[e] || f[e] != f[e]) {
          best[e] = f[e];
          bidx[e] = k;
        }
      }
    }
    pooled[i] = pack8(best);
    uint2 id;
    id.x = bidx[0] | (bidx[1] << 8) | (bidx[2] << 16) | (bidx[3] << 24);
    id.y = bidx[4] | (bidx[5] << 8) | (bidx[6] << 16) | (bidx[7] << 24);
    idx[i] = id;
  }
}
cudaError_t bn_apply_pool(const void* y, void* z, void* pooled, void* idx, const float* scale,
                          const float* shift, int N, int H, int W, int C, cudaStream_t st) {
  const long long total = static_cast<long long>(N) * (H / 2) * (W / 2) * (C / 8);
  if (scale != nullptr)
    launch_k(bn_apply_pool_kernel<true>, dim3(grid_for(total, 256 * 2, 16)), dim3(256), 0, st, 
        static_cast<const uint4*>(y), static_cast<uint4*>(z), static_cast<uint4*>(pooled),
        static_cast<uint2*>(idx), scale, shift, N, H, W, C / 8);
  else
    launch_k(bn_apply_pool_kernel<false>, dim3(grid_for(total, 256 * 2, 16)), dim3(256), 0, st, 
        static_cast<const uint4*>(y), nullptr, static_cast<uint4*>(pooled), static_cast<uint2*>(idx),
        nullptr, nullptr, N, H, W, C / 8);
  return cudaGetLastError();
}

// dIn[n, 2ho+i, 2wo+j, c] = (idx == 2i+j ? dPooled : 0) + (skip ? skip[...] : 0)
__global__ void maxpool_bwd_add_kernel(const uint4* __restrict__ dpooled, const uint2* __restrict__ idx,
                                       const uint4* __restrict__ skip, uint4* __restrict__ din, int N,
                                       int H, int W, int CV) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(N) * Ho * Wo * CV;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % CV);
    long long r = i / CV;
    const int wo = static_cast<int>(r % Wo);
    r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    const long long n = r / Ho;
    float g[8];
    unpack8(ldg_stream(dpooled + i), g);
    const uint2 id = idx[i];
    uint32_t bi[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      bi[e] = (id.x >> (8 * e)) & 0xFF;
      bi[4 + e] = (id.y >> (8 * e)) & 0xFF;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long pix = (n * H + 2 * ho + (k >> 1)) * W + 2 * wo + (k & 1);
      float f[8];
      if (skip != nullptr) {
        unpack8(ldg_stream(skip + pix * CV + cv), f);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (bi[e] == static_cast<uint32_t>(k)) f[e] += g[e];
      din[pix * CV + cv] = pack8(f);
    }
  }
}
cudaError_t maxpool_bwd_add(const void* dpooled, const void* idx, const void* skip, void* din, int N,
                            int H, int W, int C, cudaStream_t st) {
  const long long total = static_cast<long long>(N) * (H / 2) * (W / 2) * (C / 8);
  launch_k(maxpool_bwd_add_kernel, dim3(grid_for(total, 256 * 2, 16)), dim3(256), 0, st, 
      static_cast<const uint4*>(dpooled), static_cast<const uint2*>(idx), static_cast<const uint4*>(skip),
      static_cast<uint4*>(din), N, H, W, C / 8);
  return cudaGetLastError();
}

// The same, fused with the BatchNorm-backward reductions of the layer whose output was pooled (the gradient `din` it
// writes IS that layer's dz): S1 += sum din, S2 += sum din * y per channel, on the bf16 values it stores, so the separate
// bn_bwd_reduce pass over (dz, y) becomes one extra read of y here (2 instead of 4 bytes per element, one launch less).
// 256 threads, 256 % CV == 0: a thread keeps its vector column cv for the whole grid-stride loop.
__global__ void __launch_bounds__(256)
    maxpool_bwd_add_reduce_kernel(const uint4* __restrict__ dpooled, const uint2* __restrict__ idx,
                                  const uint4* __restrict__ skip, const uint4* __restrict__ y, uint4* __restrict__ din,
                                  double* __restrict__ s1, double* __restrict__ s2, int N, int H, int W, int CV) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float red[];  // [rows][CV * 16]
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(N) * Ho * Wo * CV;
  float acc[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % CV);
    long long r = i / CV;
    const int wo = static_cast<int>(r % Wo);
    r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    const long long n = r / Ho;
    float g[8];
    unpack8(ldg_stream(dpooled + i), g);
    const uint2 id = idx[i];
    uint32_t bi[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      bi[e] = (id.x >> (8 * e)) & 0xFF;
      bi[4 + e] = (id.y >> (8 * e)) & 0xFF;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long pix = (n * H + 2 * ho + (k >> 1)) * W + 2 * wo + (k & 1);
      float f[8], yv[8];
      if (skip != nullptr) {
        unpack8(ldg_stream(skip + pix * CV + cv), f);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = 0.f;
      }
      unpack8(ldg_stream(y + pix * CV + cv), yv);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (bi[e] == static_cast<uint32_t>(k)) f[e] += g[e];
      const uint4 pk = pack8(f);
      din[pix * CV + cv] = pk;
      unpack8(pk, f);  // the reductions see exactly what a later pass over the stored tensor would
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        acc[e] += f[e];
        acc[8 + e] = fmaf(f[e], yv[e], acc[8 + e]);
      }
    }
  }
  const int rows = blockDim.x / CV;
  const int cv = threadIdx.x % CV;
  const int rr = threadIdx.x / CV;
  const int width = CV * 16;
#pragma unroll
  for (int e = 0; e < 16; ++e) red[rr * width + (e / 8) * CV * 8 + cv * 8 + (e % 8)] = acc[e];
  __syncthreads();
  for (int i = threadIdx.x; i < width; i += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < rows; ++q) s += red[q * width + i];
    const int which = i / (CV * 8), c = i % (CV * 8);
    atomicAdd((which == 0 ? s1 : s2) + c, static_cast<double>(s));
  }
}
cudaError_t maxpool_bwd_add_reduce(const void* dpooled, const void* idx, const void* skip, const void* y, void* din,
                                   double* s1, double* s2, int N, int H, int W, int C, cudaStream_t st) {
  const int CV = C / 8;
  if (CV < 1 || 256 % CV != 0) return cudaErrorInvalidValue;
  const long long total = static_cast<long long>(N) * (H / 2) * (W / 2) * CV;
  const size_t smem = static_cast<size_t>(256 / CV) * CV * 16 * sizeof(float);
  launch_k(maxpool_bwd_add_reduce_kernel, dim3(grid_for(total, 256 * 2, 8)), dim3(256), smem, st,
           static_cast<const uint4*>(dpooled), static_cast<const uint2*>(idx), static_cast<const uint4*>(skip),
           static_cast<const uint4*>(y), static_cast<uint4*>(din), s1, s2, N, H, W, CV);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// per-channel reductions over NHWC bf16.  Block = 256 threads = (256/CV) pixel rows x CV vector
// columns (CV = C/8 <= 256); each thread keeps 8 (or 16) fp32 partials, rows are folded through
// shared memory, one fp64 atomic per channel per block.
template <int NACC, typename Body>
__device__ __forceinline__ void channel_reduce(long long P, int CV, double* out0, double* out1,
                                               Body body) {
  extern __shared__ float red[];  // [rows][CV*8*NACC]
  const int rows = blockDim.x / CV;
  const int cv = threadIdx.x % CV;
  const int r = threadIdx.x / CV;
  float acc[8 * NACC];
#pragma unroll
  for (int e = 0; e < 8 * NACC; ++e) acc[e] = 0.f;
  if (r < rows) {
    for (long long p = blockIdx.x * static_cast<long long>(rows) + r; p < P;
         p += static_cast<long long>(gridDim.x) * rows)
      body(p, cv, acc);
  }
  const int width = CV * 8 * NACC;
  if (r < rows) {
#pragma unroll
    for (int e = 0; e < 8 * NACC; ++e) red[r * width + (e / 8) * CV * 8 + cv * 8 + (e % 8)] = acc[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < width; i += blockDim.x) {
    float s = 0.f;
    for (int rr = 0; rr < rows; ++rr) s += red[rr * width + i];
    const int which = i / (CV * 8), c = i % (CV * 8);
    atomicAdd((which == 0 ? out0 : out1) + c, static_cast<double>(s));
  }
}

__global__ void bn_stats_kernel(const uint4* __restrict__ y, double* __restrict__ sum,
                                double* __restrict__ sq, long long P, int CV) {
  pdl_launch_dependents();
  pdl_wait();
  channel_reduce<2>(P, CV, sum, sq, [&](long long p, int cv, float* acc) {
    float f[8];
    unpack8(ldg_stream(y + p * CV + cv), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[e] += f[e];
      acc[8 + e] = fmaf(f[e], f[e], acc[8 + e]);
    }
  });
}
cudaError_t bn_stats(const void* y, double* sum, double* sq, long long P, int C, cudaStream_t st) {
  const int CV = C / 8;
  if (CV > 256 || CV < 1) return cudaErrorInvalidValue;
  const int rows = 256 / CV;
  const size_t smem = static_cast<size_t>(rows) * CV * 16 * sizeof(float);
  launch_k(bn_stats_kernel, dim3(grid_for(P, rows * 8, 8)), dim3(256), smem, st, static_cast<const uint4*>(y), sum, sq, P, CV);
  return cudaGetLastError();
}

// S1 = sum dz, S2 = sum dz*y  (per channel).  Same thread mapping as channel_reduce, two pixel rows in flight.
__global__ void __launch_bounds__(256, 4)
    bn_bwd_reduce_kernel(const uint4* __restrict__ dz, const uint4* __restrict__ y, double* __restrict__ s1,
                         double* __restrict__ s2, long long P, int CV) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float red[];
  const int rows = blockDim.x / CV;
  const int cv = threadIdx.x % CV;
  const int r = threadIdx.x / CV;
  float acc[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.f;
  // descending pixel order: the producer of dz (a dgrad epilogue) wrote it in ascending tile order, so the tail
  // is what is still in L2; bn_relu_bwd_apply then walks ascending and finds the head this kernel read last
  const long long stride = static_cast<long long>(gridDim.x) * rows;
  for (long long q = blockIdx.x * static_cast<long long>(rows) + r; r < rows && q < P; q += 2 * stride) {
    const bool two = q + stride < P;
    const long long p = P - 1 - q - (two ? stride : 0);
    const uint4 g0 = ldg_stream(dz + p * CV + cv), f0 = ldg_stream(y + p * CV + cv);
    uint4 g1 = make_uint4(0, 0, 0, 0), f1 = g1;
    if (two) {
      g1 = ldg_stream(dz + (p + stride) * CV + cv);
      f1 = ldg_stream(y + (p + stride) * CV + cv);
    }
    float g[8], f[8];
    unpack8(g0, g);
    unpack8(f0, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[e] += g[e];
      acc[8 + e] = fmaf(g[e], f[e], acc[8 + e]);
    }
    unpack8(g1, g);
    unpack8(f1, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[e] += g[e];
      acc[8 + e] = fmaf(g[e], f[e], acc[8 + e]);
    }
  }
  const int width = CV * 16;
  if (r < rows) {
#pragma unroll
    for (int e = 0; e < 16; ++e) red[r * width + (e / 8) * CV * 8 + cv * 8 + (e % 8)] = acc[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < width; i += blockDim.x) {
    float s = 0.f;
    for (int rr = 0; rr < rows; ++rr) s += red[rr * width + i];
    atomicAdd((i < CV * 8 ? s1 : s2) + (i % (CV * 8)), static_cast<double>(s));
  }
}
cudaError_t bn_bwd_reduce(const void* dz, const void* y, double* s1, double* s2, long long P, int C,
                          cudaStream_t st) {
  const int CV = C / 8;
  if (CV > 256 || CV < 1) return cudaErrorInvalidValue;
  const int rows = 256 / CV;
  const size_t smem = static_cast<size_t>(rows) * CV * 16 * sizeof(float);
  launch_k(bn_bwd_reduce_kernel, dim3(grid_for(P, rows * 8, 4)), dim3(256), smem, st, 
      static_cast<const uint4*>(dz), static_cast<const uint4*>(y), s1, s2, P, CV);
  return cudaGetLastError();
}

// dgamma = invstd*(S2 - mean*S1); dbeta = S1; coefficients of dy = kA*dz + kB*y + kC
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ s1, const double* __restrict__ s2,
                                       const float* __restrict__ gamma, const float* __restrict__ mean,
                                       const float* __restrict__ invstd, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ kA,
                                       float* __restrict__ kB, float* __restrict__ kC, int C, double count,
                                       int training, int accumulate) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = mean[c], is = invstd[c], g = gamma[c];
  const double dg = is * (s2[c] - mu * s1[c]);
  const double db = s1[c];
  if (accumulate) {
    dgamma[c] += static_cast<float>(dg);
    dbeta[c] += static_cast<float>(db);
  } else {
    dgamma[c] = static_cast<float>(dg);
    dbeta[c] = static_cast<float>(db);
  }
  kA[c] = static_cast<float>(g * is);
  if (training) {
    kB[c] = static_cast<float>(-g * is * is * dg / count);
    kC[c] = static_cast<float>(-g * is * db / count + g * is * is * mu * dg / count);
  } else {
    kB[c] = 0.f;
    kC[c] = 0.f;
  }
}
cudaError_t bn_bwd_finalize(const double* s1, const double* s2, const float* gamma, const float* mean,
                            const float* invstd, float* dgamma, float* dbeta, float* kA, float* kB,
                            float* kC, int C, double count, int training, int accumulate,
                            cudaStream_t st) {
  launch_k(bn_bwd_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, st, s1, s2, gamma, mean, invstd, dgamma, dbeta, kA,
                                                          kB, kC, C, count, training, accumulate);
  return cudaGetLastError();
}

// dpre = (y > 0) ? kA*dz + kB*y + kC : 0 ; dbias[c] += sum dpre.  Coefficients hoisted, two rows in flight.
__global__ void __launch_bounds__(256, 4)
    bn_relu_bwd_apply_kernel(const uint4* __restrict__ dz, const uint4* __restrict__ y, uint4* __restrict__ dpre,
                             const float* __restrict__ kA, const float* __restrict__ kB,
                             const float* __restrict__ kC, double* __restrict__ dbias, long long P, int CV) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float red[];
  const int rows = blockDim.x / CV;
  const int cv = threadIdx.x % CV;
  const int r = threadIdx.x / CV;
  float ka[8], kb[8], kc[8], acc[8];
  {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(kA) + cv * 2), a1 = __ldg(reinterpret_cast<const float4*>(kA) + cv * 2 + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(kB) + cv * 2), b1 = __ldg(reinterpret_cast<const float4*>(kB) + cv * 2 + 1);
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(kC) + cv * 2), c1 = __ldg(reinterpret_cast<const float4*>(kC) + cv * 2 + 1);
    ka[0] = a0.x; ka[1] = a0.y; ka[2] = a0.z; ka[3] = a0.w; ka[4] = a1.x; ka[5] = a1.y; ka[6] = a1.z; ka[7] = a1.w;
    kb[0] = b0.x; kb[1] = b0.y; kb[2] = b0.z; kb[3] = b0.w; kb[4] = b1.x; kb[5] = b1.y; kb[6] = b1.z; kb[7] = b1.w;
    kc[0] = c0.x; kc[1] = c0.y; kc[2] = c0.z; kc[3] = c0.w; kc[4] = c1.x; kc[5] = c1.y; kc[6] = c1.z; kc[7] = c1.w;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  // ascending pixel order; bn_bwd_reduce walked the same two tensors in DESCENDING order just before, so the
  // head of dz / y is what is still resident in the 126 MB L2
  const long long stride = static_cast<long long>(gridDim.x) * rows;
  for (long long p = blockIdx.x * static_cast<long long>(rows) + r; r < rows && p < P; p += 2 * stride) {
    const bool two = p + stride < P;
    const uint4 g0 = ldg_stream(dz + p * CV + cv), f0 = ldg_stream(y + p * CV + cv);
    uint4 g1 = make_uint4(0, 0, 0, 0), f1 = g1;
    if (two) {
      g1 = ldg_stream(dz + (p + stride) * CV + cv);
      f1 = ldg_stream(y + (p + stride) * CV + cv);
    }
    float g[8], f[8], o[8];
    unpack8(g0, g);
    unpack8(f0, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float d = fmaf(ka[e], g[e], fmaf(kb[e], f[e], kc[e]));
      o[e] = f[e] > 0.f ? d : 0.f;
      acc[e] += o[e];
    }
    dpre[p * CV + cv] = pack8(o);
    if (two) {
      unpack8(g1, g);
      unpack8(f1, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = fmaf(ka[e], g[e], fmaf(kb[e], f[e], kc[e]));
        o[e] = f[e] > 0.f ? d : 0.f;
        acc[e] += o[e];
      }
      dpre[(p + stride) * CV + cv] = pack8(o);
    }
  }
  const int width = CV * 8;
  if (r < rows) {
#pragma unroll
    for (int e = 0; e < 8; ++e) red[r * width + cv * 8 + e] = acc[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < width; i += blockDim.x) {
    float s = 0.f;
    for (int rr = 0; rr < rows; ++rr) s += red[rr * width + i];
    atomicAdd(dbias + i, static_cast<double>(s));
  }
}
cudaError_t bn_relu_bwd_apply(const void* dz, const void* y, void* dpre, const float* kA, const float* kB,
                              const float* kC, double* dbias, long long P, int C, cudaStream_t st) {
  const int CV = C / 8;
  if (CV > 256 || CV < 1) return cudaErrorInvalidValue;
  const int rows = 256 / CV;
  const size_t smem = static_cast<size_t>(rows) * CV * 8 * sizeof(float);
  launch_k(bn_relu_bwd_apply_kernel, dim3(grid_for(P, rows * 8, 4)), dim3(256), smem, st, 
      static_cast<const uint4*>(dz), static_cast<const uint4*>(y), static_cast<uint4*>(dpre), kA, kB, kC,
      dbias, P, CV);
  return cudaGetLastError();
}

// per-channel sum of a bf16 NHWC tensor (bias gradient of ConvTranspose2d / conv1x1)
__global__ void channel_sum_kernel(const uint4* __restrict__ g, double* __restrict__ out, long long P,
                                   int CV) {
  pdl_launch_dependents();
  pdl_wait();
  channel_reduce<1>(P, CV, out, out, [&](long long p, int cv, float* acc) {
    float f[8];
    unpack8(ldg_stream(g + p * CV + cv), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += f[e];
  });
}
cudaError_t channel_sum(const void* g, double* out, long long P, int C, cudaStream_t st) {
  const int CV = C / 8;
  if (CV > 256 || CV < 1) return cudaErrorInvalidValue;
  const int rows = 256 / CV;
  const size_t smem = static_cast<size_t>(rows) * CV * 8 * sizeof(float);
  launch_k(channel_sum_kernel, dim3(grid_for(P, rows * 8, 8)), dim3(256), smem, st, static_cast<const uint4*>(g), out, P, CV);
  return cudaGetLastError();
}

// dst[i] = (accumulate ? dst[i] : 0) + alpha*src[i]   (fp64 accumulators -> fp32 .grad)
__global__ void f64_to_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, int n, int ld_group,
                                  int groups, float alpha, int accumulate) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int g = 0; g < groups; ++g) s += src[g * ld_group + i];
  const float v = alpha * static_cast<float>(s);
  dst[i] = accumulate ? dst[i] + v : v;
}
cudaError_t f64_to_f32(const double* src, float* dst, int n, int ld_group, int groups, float alpha,
                       int accumulate, cudaStream_t st) {
  launch_k(f64_to_f32_kernel, dim3((n + 127) / 128), dim3(128), 0, st, src, dst, n, ld_group, groups, alpha, accumulate);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// fused softmax cross-entropy (+ temperature-KL distillation) forward + backward over fp32 logits.
//   loss_acc[0] += sum_p -log softmax(z_p)[y_p]
//   loss_acc[1] += sum_p sum_{c<Cold} p0_c (log p0_c - log q_c),  p0 = softmax(zold/T), q = softmax(z[:Cold]/T)
//   dlogits[p][c] = gscale*((softmax(z)_c - [c==y]) + [c<Cold]*lambda*T*(q_c - p0_c)), bf16, row pitch ldd
constexpr int kLossPix = 256;
__global__ void __launch_bounds__(kLossPix)
    ce_kd_loss_kernel(const float* __restrict__ logits, const float* __restrict__ old_logits,
                      const long long* __restrict__ labels, long long P, int C, int Cold, float T,
                      float lambda, float gscale, __nv_bfloat16* __restrict__ dlogits, int ldd,
                      double* __restrict__ loss_acc, int* __restrict__ err_flag) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];  // [256*C] new logits, [256*Cold] old logits
  float* sz = sm;
  float* so = sm + kLossPix * C;
  __shared__ float wsum[2][kLossPix / 32];
  float ce_local = 0.f, kd_local = 0.f;
  const float invT = 1.f / T;
  for (long long base = static_cast<long long>(blockIdx.x) * kLossPix; base < P;
       base += static_cast<long long>(gridDim.x) * kLossPix) {
    const int npx = static_cast<int>(min(static_cast<long long>(kLossPix), P - base));
    __syncthreads();
    for (int i = threadIdx.x; i < npx * C; i += kLossPix) sz[i] = logits[base * C + i];
    if (old_logits != nullptr)
      for (int i = threadIdx.x; i < npx * Cold; i += kLossPix) so[i] = old_logits[base * Cold + i];
    __syncthreads();
    if (threadIdx.x < npx) {
      const float* z = sz + threadIdx.x * C;
      const long long p = base + threadIdx.x;
      const long long y = labels[p];
      float mx = -CUDART_INF_F;
      for (int c = 0; c < C; ++c) mx = fmaxf(mx, z[c]);
      float se = 0.f;
      for (int c = 0; c < C; ++c) se += __expf(z[c] - mx);
      const float lse = mx + __logf(se);
      const float inv_se = 1.f / se;
      bool yok = y >= 0 && y < C;
      if (!yok && err_flag != nullptr) *err_flag = 1;
      if (yok) ce_local += lse - z[y];
      // distillation terms
      float mq = 0.f, sq = 1.f, mo = 0.f, so_ = 1.f, lq = 0.f, lo = 0.f;
      const float* zo = so + threadIdx.x * Cold;
      if (old_logits != nullptr) {
        mq = -CUDART_INF_F;
        mo = -CUDART_INF_F;
        for (int c = 0; c < Cold; ++c) {
          mq = fmaxf(mq, z[c] * invT);
          mo = fmaxf(mo, zo[c] * invT);
        }
        sq = 0.f;
        so_ = 0.f;
        for (int c = 0; c < Cold; ++c) {
          sq += __expf(z[c] * invT - mq);
          so_ += __expf(zo[c] * invT - mo);
        }
        lq = mq + __logf(sq);
        lo = mo + __logf(so_);
      }
      __nv_bfloat16* d = dlogits + p * ldd;
      float kd = 0.f;
      for (int c0 = 0; c0 < ldd; c0 += 8) {
        float g[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = c0 + e;
          float v = 0.f;
          if (c < C) {
            v = yok ? (__expf(z[c] - mx) * inv_se - (c == y ? 1.f : 0.f)) : 0.f;
            if (old_logits != nullptr && c < Cold) {
              const float logq = z[c] * invT - lq;
              const float logp0 = zo[c] * invT - lo;
              const float p0 = __expf(logp0);
              kd += p0 * (logp0 - logq);
              v += lambda * T * (__expf(logq) - p0);
            }
            v *= gscale;
          }
          g[e] = v;
        }
        *reinterpret_cast<uint4*>(d + c0) = pack8(g);
      }
      kd_local += kd;
    }
  }
  // block reduce the two partial sums
  for (int o = 16; o > 0; o >>= 1) {
    ce_local += __shfl_xor_sync(0xffffffffu, ce_local, o);
    kd_local += __shfl_xor_sync(0xffffffffu, kd_local, o);
  }
  if ((threadIdx.x & 31) == 0) {
    wsum[0][threadIdx.x >> 5] = ce_local;
    wsum[1][threadIdx.x >> 5] = kd_local;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < kLossPix / 32; ++w) {
      a += wsum[0][w];
      b += wsum[1][w];
    }
    atomicAdd(loss_acc, a);
    atomicAdd(loss_acc + 1, b);
  }
}
cudaError_t ce_kd_loss(const float* logits, const float* old_logits, const long long* labels, long long P,
                       int C, int Cold, float T, float lambda, float gscale, void* dlogits, int ldd,
                       double* loss_acc, int* err_flag, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(kLossPix) * (C + (old_logits ? Cold : 0)) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaFuncSetAttribute(ce_kd_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  launch_k(ce_kd_loss_kernel, dim3(grid_for(P, kLossPix, 4)), dim3(kLossPix), smem, st, 
      logits, old_logits, labels, P, C, Cold, T, lambda, gscale, static_cast<__nv_bfloat16*>(dlogits), ldd,
      loss_acc, err_flag);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// confusion matrix: shared-memory privatised histogram (one copy per warp), int64 global bins.
// Reference: metrics._fast_conf_matrix (metrics.py:32-38): rows = target, cols = prediction, targets
// outside [0, nc) are skipped; a prediction outside [0, nc) on a kept target raises in the
// reference (bincount/reshape), here it sets *err_flag and is not counted.
constexpr int kHistThreads = 256;
__device__ __forceinline__ void hist_flush(unsigned int* sh, int nbins, int copies,
                                           unsigned long long* __restrict__ conf) {
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += blockDim.x) {
    unsigned long long s = 0;
    for (int w = 0; w < copies; ++w) s += sh[w * nbins + b];
    if (s) atomicAdd(conf + b, s);
  }
}
__global__ void __launch_bounds__(kHistThreads)
    confusion_kernel(const long long* __restrict__ target, const long long* __restrict__ pred,
                     long long n, int nc, unsigned long long* __restrict__ conf, int* __restrict__ err_flag) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ unsigned int sh[];  // [copies][nc*nc]
  const int nbins = nc * nc;
  const int copies = kHistThreads / 32;
  for (int i = threadIdx.x; i < copies * nbins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  unsigned int* mine = sh + (threadIdx.x >> 5) * nbins;
  // two int64 per 16-byte vector when both inputs are 16-byte aligned, scalar loads otherwise
  const bool vec = ((reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(pred)) & 15) == 0;
  const long long nvec = vec ? n / 2 : 0;
  const longlong2* t2 = reinterpret_cast<const longlong2*>(target);
  const longlong2* p2 = reinterpret_cast<const longlong2*>(pred);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const longlong2 t = t2[i], p = p2[i];
    if (t.x >= 0 && t.x < nc) {
      if (p.x >= 0 && p.x < nc) atomicAdd(mine + t.x * nc + p.x, 1u);
      else if (err_flag) *err_flag = 1;
    }
    if (t.y >= 0 && t.y < nc) {
      if (p.y >= 0 && p.y < nc) atomicAdd(mine + t.y * nc + p.y, 1u);
      else if (err_flag) *err_flag = 1;
    }
  }
  for (long long i = 2 * nvec + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long t = target[i], p = pred[i];
    if (t >= 0 && t < nc) {
      if (p >= 0 && p < nc) atomicAdd(mine + t * nc + p, 1u);
      else if (err_flag) *err_flag = 1;
    }
  }
  hist_flush(sh, nbins, copies, conf);
}
cudaError_t confusion_matrix(const long long* target, const long long* pred, long long n, int nc,
                             long long* conf, int* err_flag, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(kHistThreads / 32) * nc * nc * sizeof(unsigned int);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  // each thread must see < 2^32 items per bin copy: guaranteed for n < 2^32 per block pass
  launch_k(confusion_kernel, dim3(grid_for(n / 2 + 1, kHistThreads * 8, 4)), dim3(kHistThreads), smem, st, 
      target, pred, n, nc, reinterpret_cast<unsigned long long*>(conf), err_flag);
  return cudaGetLastError();
}

// argmax over C fp32 logits per pixel (first maximum, like torch.argmax / torch.max on CPU and CUDA)
// fused with the confusion histogram and the correct-pixel count (trainer.py:183-184, 279-280).
__global__ void __launch_bounds__(kHistThreads)
    argmax_confusion_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                            long long P, int C, int nc, long long* __restrict__ pred_out,
                            unsigned long long* __restrict__ conf, unsigned long long* __restrict__ correct) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ unsigned int sh[];  // [copies][nc*nc] then [256*C] floats
  const int nbins = nc * nc;
  const int copies = kHistThreads / 32;
  float* sz = reinterpret_cast<float*>(sh + copies * nbins);
  for (int i = threadIdx.x; i < copies * nbins; i += blockDim.x) sh[i] = 0;
  unsigned int* mine = sh + (threadIdx.x >> 5) * nbins;
  unsigned int ok = 0;
  for (long long base = static_cast<long long>(blockIdx.x) * kHistThreads; base < P;
       base += static_cast<long long>(gridDim.x) * kHistThreads) {
    const int npx = static_cast<int>(min(static_cast<long long>(kHistThreads), P - base));
    __syncthreads();
    for (int i = threadIdx.x; i < npx * C; i += kHistThreads) sz[i] = logits[base * C + i];
    __syncthreads();
    if (threadIdx.x < npx) {
      const float* z = sz + threadIdx.x * C;
      float best = z[0];
      int bi = 0;
      for (int c = 1; c < C; ++c) {
        const float v = z[c];
        if (v > best || (v != v && best == best)) {
          best = v;
          bi = c;
        }
      }
      const long long p = base + threadIdx.x;
      const long long t = labels[p];
      if (pred_out != nullptr) pred_out[p] = bi;
      if (t == bi) ++ok;
      if (conf != nullptr && t >= 0 && t < nc) atomicAdd(mine + t * nc + bi, 1u);
    }
  }
  for (int o = 16; o > 0; o >>= 1) ok += __shfl_xor_sync(0xffffffffu, ok, o);
  if ((threadIdx.x & 31) == 0 && ok && correct != nullptr) atomicAdd(correct, static_cast<unsigned long long>(ok));
  if (conf != nullptr) hist_flush(sh, nbins, copies, conf);
}
cudaError_t argmax_confusion(const float* logits, const long long* labels, long long P, int C, int nc,
                             long long* pred_out, long long* conf, long long* correct, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(kHistThreads / 32) * nc * nc * sizeof(unsigned int) +
                      static_cast<size_t>(kHistThreads) * C * sizeof(float);
  if (smem > 96 * 1024) return cudaErrorInvalidValue;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaFuncSetAttribute(argmax_confusion_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  }
  launch_k(argmax_confusion_kernel, dim3(grid_for(P, kHistThreads * 4, 4)), dim3(kHistThreads), smem, st, 
      logits, labels, P, C, nc, pred_out, reinterpret_cast<unsigned long long*>(conf),
      reinterpret_cast<unsigned long long*>(correct));
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// multi-tensor Adam (torch.optim.Adam semantics, amsgrad=False, weight_decay=0, maximize=False):
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void adam_kernel(const AdamTensor* __restrict__ tensors, const int2* __restrict__ blocks,
                            int chunk, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt,
                            float gscale, const float* __restrict__ hyper) {
  pdl_launch_dependents();
  pdl_wait();
  if (hyper != nullptr) {  // {lr, bc1, bc2_sqrt, gscale} in device memory (CUDA-graph replay)
    lr = hyper[0];
    bc1 = hyper[1];
    bc2_sqrt = hyper[2];
    gscale = hyper[3];
  }
  const int2 blk = blocks[blockIdx.x];
  const AdamTensor t = tensors[blk.x];
  const long long begin = static_cast<long long>(blk.y) * chunk;
  const long long end = min(t.numel, begin + chunk);
  const float step_size = lr / bc1;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                        reinterpret_cast<uintptr_t>(t.m) | reinterpret_cast<uintptr_t>(t.v)) & 15) == 0;
  if (vec_ok) {
    const long long vend = begin + ((end - begin) / 4) * 4;
    for (long long i = begin + threadIdx.x * 4; i < vend; i += blockDim.x * 4) {
      float4 p = *reinterpret_cast<float4*>(t.p + i);
      const float4 g4 = *reinterpret_cast<const float4*>(t.g + i);
      float4 m = *reinterpret_cast<float4*>(t.m + i);
      float4 v = *reinterpret_cast<float4*>(t.v + i);
      float* pp = &p.x; const float* gg = &g4.x; float* mm = &m.x; float* vv = &v.x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float g = gg[e] * gscale;
        mm[e] = b1 * mm[e] + (1.f - b1) * g;
        vv[e] = b2 * vv[e] + (1.f - b2) * g * g;
        const float denom = sqrtf(vv[e]) / bc2_sqrt + eps;
        pp[e] -= step_size * (mm[e] / denom);
      }
      *reinterpret_cast<float4*>(t.p + i) = p;
      *reinterpret_cast<float4*>(t.m + i) = m;
      *reinterpret_cast<float4*>(t.v + i) = v;
    }
    for (long long i = vend + threadIdx.x; i < end; i += blockDim.x) {
      const float g = t.g[i] * gscale;
      const float m = b1 * t.m[i] + (1.f - b1) * g;
      const float v = b2 * t.v[i] + (1.f - b2) * g * g;
      t.m[i] = m; t.v[i] = v;
      t.p[i] -= step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
    }
  } else {
    for (long long i = begin + threadIdx.x; i < end; i += blockDim.x) {
      const float g = t.g[i] * gscale;
      const float m = b1 * t.m[i] + (1.f - b1) * g;
      const float v = b2 * t.v[i] + (1.f - b2) * g * g;
      t.m[i] = m; t.v[i] = v;
      t.p[i] -= step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
    }
  }
}
cudaError_t adam_multi_tensor(const AdamTensor* tensors, const void* blocks, int nblocks, int chunk,
                              float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt,
                              float gscale, const float* hyper, cudaStream_t st) {
  if (nblocks <= 0) return cudaSuccess;
  launch_k(adam_kernel, dim3(nblocks), dim3(256), 0, st, tensors, static_cast<const int2*>(blocks), chunk, lr, b1, b2, eps,
                                       bc1, bc2_sqrt, gscale, hyper);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// table-driven batched variants: ONE launch packs / unpacks / converts every layer of the model
// (per-layer launches of these tiny kernels cost more in launch latency than in bandwidth).
// A job table is a device array of int64[16] rows; one thread block = one 32x32(xT) tile of one job.
__device__ __forceinline__ int find_job(const long long* __restrict__ jobs, int njobs, int tile, int tile_col) {
  int j = 0;
  while (j + 1 < njobs && jobs[(j + 1) * 16 + tile_col] <= tile) ++j;
  return j;
}

// full 32x32 tile with compile-time T: every global load of a thread is independent and issued before the first
// use (4*T loads in flight per thread), and the bf16 outputs leave as 16-byte stores (8 rows x 64 B per warp
// instruction instead of one 64-byte row).  Shared-memory indices are conflict-free (odd pitch).
template <int T>
__device__ __forceinline__ void pack_tile_fast(float* tile, const float* __restrict__ src,
                                               __nv_bfloat16* __restrict__ outAB, __nv_bfloat16* __restrict__ outBA,
                                               int B, int ldA, int ldB, int ldB2, int ldA2, int rev, int a0, int b0,
                                               int warp, int lane) {
  constexpr int pitch = 32 * T + 1;
  float v[4][T];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float* row = src + (static_cast<long long>(a0 + warp + 8 * r) * B + b0) * T;
#pragma unroll
    for (int i = 0; i < T; ++i) v[r][i] = row[lane + 32 * i];
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int i = 0; i < T; ++i) tile[(warp + 8 * r) * pitch + lane + 32 * i] = v[r][i];
  __syncthreads();
  const int rr = lane >> 2, c = lane & 3;
  if (outAB != nullptr) {
    for (int item = warp; item < 4 * T; item += 8) {
      const int t = item >> 2, ar = (item & 3) * 8 + rr;
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = tile[ar * pitch + (8 * c + e) * T + t];
      *reinterpret_cast<uint4*>(outAB + (static_cast<long long>(t) * ldA + a0 + ar) * ldB + b0 + 8 * c) = pack8(f);
    }
  }
  if (outBA != nullptr) {
    for (int item = warp; item < 4 * T; item += 8) {
      const int t = item >> 2, br = (item & 3) * 8 + rr;
      const int tt = rev ? T - 1 - t : t;
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = tile[(8 * c + e) * pitch + br * T + t];
      *reinterpret_cast<uint4*>(outBA + (static_cast<long long>(tt) * ldB2 + b0 + br) * ldA2 + a0 + 8 * c) = pack8(f);
    }
  }
}

// row: {src, outAB, outBA, A, B, T, ldA, ldB, ldB2, ldA2, rev, tile0, tiles_b, -, -, -}
__global__ void __launch_bounds__(256) pack_w_multi_kernel(const long long* __restrict__ jobs, int njobs) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float tile[];  // [32][32*T + 1]
  const int j = find_job(jobs, njobs, blockIdx.x, 11);
  const long long* J = jobs + j * 16;
  const float* __restrict__ src = reinterpret_cast<const float*>(J[0]);
  __nv_bfloat16* __restrict__ outAB = reinterpret_cast<__nv_bfloat16*>(J[1]);
  __nv_bfloat16* __restrict__ outBA = reinterpret_cast<__nv_bfloat16*>(J[2]);
  const int A = static_cast<int>(J[3]), B = static_cast<int>(J[4]), T = static_cast<int>(J[5]);
  const int ldA = static_cast<int>(J[6]), ldB = static_cast<int>(J[7]);
  const int ldB2 = static_cast<int>(J[8]), ldA2 = static_cast<int>(J[9]), rev = static_cast<int>(J[10]);
  const int local = blockIdx.x - static_cast<int>(J[11]);
  const int tiles_b = static_cast<int>(J[12]);
  const int a0 = (local / tiles_b) * 32, b0 = (local % tiles_b) * 32;
  const int na = min(32, A - a0), nb = min(32, B - b0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pitch = 32 * T + 1;
  const bool vec_ok = na == 32 && nb == 32 && (ldB % 8) == 0 && (ldA2 % 8) == 0 &&
                      ((reinterpret_cast<uintptr_t>(outAB) | reinterpret_cast<uintptr_t>(outBA)) & 15) == 0;
  if (vec_ok && T == 9) {
    pack_tile_fast<9>(tile, src, outAB, outBA, B, ldA, ldB, ldB2, ldA2, rev, a0, b0, warp, lane);
    return;
  }
  if (vec_ok && T == 4) {
    pack_tile_fast<4>(tile, src, outAB, outBA, B, ldA, ldB, ldB2, ldA2, rev, a0, b0, warp, lane);
    return;
  }
  for (int ar = warp; ar < na; ar += 8) {
    const float* row = src + (static_cast<long long>(a0 + ar) * B + b0) * T;
    for (int rem = lane; rem < nb * T; rem += 32) tile[ar * pitch + rem] = row[rem];
  }
  __syncthreads();
  if (outAB != nullptr) {
    for (int t = 0; t < T; ++t)
      for (int ar = warp; ar < na; ar += 8)
        if (lane < nb)
          outAB[(static_cast<long long>(t) * ldA + a0 + ar) * ldB + b0 + lane] =
              __float2bfloat16_rn(tile[ar * pitch + lane * T + t]);
  }
  if (outBA != nullptr) {
    for (int t = 0; t < T; ++t) {
      const int tt = rev ? T - 1 - t : t;
      for (int br = warp; br < nb; br += 8)
        if (lane < na)
          outBA[(static_cast<long long>(tt) * ldB2 + b0 + br) * ldA2 + a0 + lane] =
              __float2bfloat16_rn(tile[lane * pitch + br * T + t]);
    }
  }
}
cudaError_t pack_w_multi(const void* jobs, int njobs, int total_tiles, int max_T, cudaStream_t st) {
  if (total_tiles <= 0) return cudaSuccess;
  const size_t smem = static_cast<size_t>(32) * (32 * max_T + 1) * sizeof(float);
  launch_k(pack_w_multi_kernel, dim3(total_tiles), dim3(256), smem, st, static_cast<const long long*>(jobs), njobs);
  return cudaGetLastError();
}

// conv3x3 case ([T][ldB][ldA] -> [A][B][T]), full tile: 4*T independent loads per thread in flight
template <int T>
__device__ __forceinline__ void unpack_tile_fast(float* tile, const float* __restrict__ D, float* __restrict__ grad,
                                                 int B, int ldA, int ldB, float alpha, int accumulate, int a0, int b0,
                                                 int warp, int lane) {
  constexpr int pitch = 32 * T + 1;
  float v[4][T];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int t = 0; t < T; ++t)
      v[r][t] = D[(static_cast<long long>(t) * ldB + b0 + warp + 8 * r) * ldA + a0 + lane];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int t = 0; t < T; ++t) tile[lane * pitch + (warp + 8 * r) * T + t] = v[r][t];
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float* row = grad + (static_cast<long long>(a0 + warp + 8 * r) * B + b0) * T;
    float old[T];
    if (accumulate) {
#pragma unroll
      for (int i = 0; i < T; ++i) old[i] = row[lane + 32 * i];
    }
#pragma unroll
    for (int i = 0; i < T; ++i) {
      const float x = alpha * tile[(warp + 8 * r) * pitch + lane + 32 * i];
      row[lane + 32 * i] = accumulate ? old[i] + x : x;
    }
  }
}

// row: {D, grad, A, B, T, ldA, ldB, alpha (double bits), accumulate, tile0, tiles_b, ...}
__global__ void __launch_bounds__(256) unpack_wgrad_multi_kernel(const long long* __restrict__ jobs, int njobs) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float tile[];
  const int j = find_job(jobs, njobs, blockIdx.x, 9);
  const long long* J = jobs + j * 16;
  const float* __restrict__ D = reinterpret_cast<const float*>(J[0]);
  float* __restrict__ grad = reinterpret_cast<float*>(J[1]);
  const int A = static_cast<int>(J[2]), B = static_cast<int>(J[3]), T = static_cast<int>(J[4]);
  const int ldA = static_cast<int>(J[5]), ldB = static_cast<int>(J[6]);
  const float alpha = static_cast<float>(__longlong_as_double(J[7]));
  const int accumulate = static_cast<int>(J[8]);
  const int local = blockIdx.x - static_cast<int>(J[9]);
  const int tiles_b = static_cast<int>(J[10]);
  const int transposed = static_cast<int>(J[11]);  // D is [T][ldB][ldA] instead of [T][ldA][ldB]
  const int a0 = (local / tiles_b) * 32, b0 = (local % tiles_b) * 32;
  const int na = min(32, A - a0), nb = min(32, B - b0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pitch = 32 * T + 1;
  const int nsplit = static_cast<int>(J[12]) > 0 ? static_cast<int>(J[12]) : 1;  // split-K partial buffers to sum
  const long long sstride = J[13];
  if (transposed && nsplit == 1 && na == 32 && nb == 32 && T == 9) {
    unpack_tile_fast<9>(tile, D, grad, B, ldA, ldB, alpha, accumulate, a0, b0, warp, lane);
    return;
  }
  if (transposed) {
    for (int t = 0; t < T; ++t)
      for (int br = warp; br < nb; br += 8)
        if (lane < na) {
          const float* src = D + (static_cast<long long>(t) * ldB + b0 + br) * ldA + a0 + lane;
          float v = 0.f;
          for (int sp = 0; sp < nsplit; ++sp) v += src[sp * sstride];  // fixed order: deterministic
          tile[lane * pitch + br * T + t] = v;
        }
  } else {
    for (int t = 0; t < T; ++t)
      for (int ar = warp; ar < na; ar += 8)
        if (lane < nb) tile[ar * pitch + lane * T + t] = D[(static_cast<long long>(t) * ldA + a0 + ar) * ldB + b0 + lane];
  }
  __syncthreads();
  for (int ar = warp; ar < na; ar += 8) {
    float* row = grad + (static_cast<long long>(a0 + ar) * B + b0) * T;
    for (int rem = lane; rem < nb * T; rem += 32) {
      const float v = alpha * tile[ar * pitch + rem];
      row[rem] = accumulate ? row[rem] + v : v;
    }
  }
}
cudaError_t unpack_wgrad_multi(const void* jobs, int njobs, int total_tiles, int max_T, cudaStream_t st) {
  if (total_tiles <= 0) return cudaSuccess;
  const size_t smem = static_cast<size_t>(32) * (32 * max_T + 1) * sizeof(float);
  launch_k(unpack_wgrad_multi_kernel, dim3(total_tiles), dim3(256), smem, st, static_cast<const long long*>(jobs), njobs);
  return cudaGetLastError();
}

// split-K partial buffers -> their sum, in place in split 0 (fixed summation order: deterministic).
// row: {base, n_vec4, nsplit, stride_vec4, block0}; one thread per float4, 8 independent loads in flight.
__global__ void __launch_bounds__(256) reduce_partials_multi_kernel(const long long* __restrict__ jobs, int njobs) {
  pdl_launch_dependents();
  pdl_wait();
  const int j = find_job(jobs, njobs, blockIdx.x, 4);
  const long long* J = jobs + j * 16;
  float4* base = reinterpret_cast<float4*>(J[0]);
  const long long nvec = J[1];
  const int nsplit = static_cast<int>(J[2]);
  const long long stride = J[3];
  const long long i = (static_cast<long long>(blockIdx.x) - J[4]) * blockDim.x + threadIdx.x;
  if (i >= nvec || nsplit <= 1) return;
  float4 acc = base[i];
  int sp = 1;
  for (; sp + 8 <= nsplit; sp += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = base[i + (sp + u) * stride];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
    }
  }
  for (; sp < nsplit; ++sp) {
    const float4 v = base[i + sp * stride];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  base[i] = acc;
}
cudaError_t reduce_partials_multi(const void* jobs, int njobs, int total_blocks, cudaStream_t st) {
  if (total_blocks <= 0) return cudaSuccess;
  launch_k(reduce_partials_multi_kernel, dim3(total_blocks), dim3(256), 0, st, static_cast<const long long*>(jobs), njobs);
  return cudaGetLastError();
}

// row: {src f64, dst f32, n, ld_group, groups, alpha (double bits), accumulate}; one block per job
__global__ void f64_to_f32_multi_kernel(const long long* __restrict__ jobs) {
  pdl_launch_dependents();
  pdl_wait();
  const long long* J = jobs + blockIdx.x * 16;
  const double* __restrict__ src = reinterpret_cast<const double*>(J[0]);
  float* __restrict__ dst = reinterpret_cast<float*>(J[1]);
  const int n = static_cast<int>(J[2]), ld = static_cast<int>(J[3]), groups = static_cast<int>(J[4]);
  const float alpha = static_cast<float>(__longlong_as_double(J[5]));
  const int accumulate = static_cast<int>(J[6]);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int g = 0; g < groups; ++g) s += src[g * ld + i];
    const float v = alpha * static_cast<float>(s);
    dst[i] = accumulate ? dst[i] + v : v;
  }
}
cudaError_t f64_to_f32_multi(const void* jobs, int njobs, cudaStream_t st) {
  if (njobs <= 0) return cudaSuccess;
  launch_k(f64_to_f32_multi_kernel, dim3(njobs), dim3(256), 0, st, static_cast<const long long*>(jobs));
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Data contract around the step: what datasets/voc.py does per pixel in Python loops.
// palette of datasets/voc.py:33-54 as 0xRRGGBB; entry 21 = void
__constant__ unsigned int c_voc_palette[22] = {
    0x000000, 0x800000, 0x008000, 0x808000, 0x000080, 0x800080, 0x008080, 0x808080, 0x400000, 0xC00000, 0x408000,
    0xC08000, 0x400080, 0xC00080, 0x408080, 0xC08080, 0x004000, 0x804000, 0x00C000, 0x80C000, 0x004080, 0xE0E0C0};

// One VOC.__getitem__ (datasets/voc.py:127-140) per blockIdx.y after decoding: Pad(10) + CenterCrop((H, W)) as a
// crop origin (top, left) in source coordinates computed on the host (zero fill outside the source), then
// ToTensor + Normalize(0.5, 0.5) (main.py:20-21) for the image and to_mask (voc.py:56-72) for the mask.
// items: int64[8] rows {img u8 [Hs][Ws][3], mask u8 [Hs][Ws][3] or 0, Hs, Ws, top, left, -, -}
__global__ void __launch_bounds__(256)
    voc_prepare_kernel(const long long* __restrict__ items, int H, int W, float* __restrict__ x,
                       long long* __restrict__ y, int* __restrict__ err_flag) {
  pdl_launch_dependents();
  pdl_wait();
  const long long* J = items + static_cast<long long>(blockIdx.y) * 8;
  const unsigned char* __restrict__ img = reinterpret_cast<const unsigned char*>(J[0]);
  const unsigned char* __restrict__ msk = reinterpret_cast<const unsigned char*>(J[1]);
  const int Hs = static_cast<int>(J[2]), Ws = static_cast<int>(J[3]);
  const int top = static_cast<int>(J[4]), left = static_cast<int>(J[5]);
  const long long hw = static_cast<long long>(H) * W;
  for (long long q = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; q < hw;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(q / W), j = static_cast<int>(q % W);
    const int si = i + top, sj = j + left;
    const bool inside = si >= 0 && si < Hs && sj >= 0 && sj < Ws;
    const long long so = (static_cast<long long>(si) * Ws + sj) * 3;
    if (img != nullptr && x != nullptr) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = inside ? static_cast<float>(img[so + c]) : 0.f;
        // ToTensor: v / 255; Normalize: (t - 0.5) / 0.5 — the reference's fp32 operations, no contraction
        const float t = __fdiv_rn(v, 255.f);
        x[(static_cast<long long>(blockIdx.y) * 3 + c) * hw + q] = __fdiv_rn(__fsub_rn(t, 0.5f), 0.5f);
      }
    }
    if (msk != nullptr && y != nullptr) {
      unsigned int key = 0;
      if (inside) key = (static_cast<unsigned int>(msk[so]) << 16) | (static_cast<unsigned int>(msk[so + 1]) << 8) | msk[so + 2];
      int label = -1;
#pragma unroll
      for (int k = 21; k >= 0; --k)
        if (key == c_voc_palette[k]) label = k;  // first match wins, as list.index
      if (label == 21) label = 0;                // void -> background (voc.py:67-68)
      if (label < 0 && err_flag != nullptr) *err_flag = 1;  // palette.index raises ValueError in the reference
      y[static_cast<long long>(blockIdx.y) * hw + q] = label;
    }
  }
}
cudaError_t voc_prepare_batch(const void* items, int B, int H, int W, float* x, long long* y, int* err_flag,
                              cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  const long long hw = static_cast<long long>(H) * W;
  int gx = static_cast<int>((hw + 255) / 256);
  if (gx > 4096) gx = 4096;
  launch_k(voc_prepare_kernel, dim3(gx, B), dim3(256), 0, st, static_cast<const long long*>(items), H, W, x, y, err_flag);
  return cudaGetLastError();
}

// datasets/voc.py:74-89 (to_rgb): class index -> palette colour, float64 [B][3][hw]; indices outside [0, 22) keep
// their own value in all three channels, exactly as the reference's repeat-then-overwrite does
__global__ void __launch_bounds__(256)
    labels_to_rgb_kernel(const long long* __restrict__ labels, long long n_images, long long hw, double* __restrict__ rgb) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = n_images * hw;
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long l = labels[p];
    const long long b = p / hw, q = p - b * hw;
    double r = static_cast<double>(l), g = r, bl = r;
    if (l >= 0 && l < 22) {
      const unsigned int c = c_voc_palette[l];
      r = static_cast<double>((c >> 16) & 255u);
      g = static_cast<double>((c >> 8) & 255u);
      bl = static_cast<double>(c & 255u);
    }
    double* o = rgb + b * 3 * hw + q;
    o[0] = r;
    o[hw] = g;
    o[2 * hw] = bl;
  }
}
cudaError_t labels_to_rgb(const long long* labels, long long n_images, long long hw, double* rgb, cudaStream_t st) {
  if (n_images * hw <= 0) return cudaSuccess;
  launch_k(labels_to_rgb_kernel, dim3(grid_for(n_images * hw, 256, 8)), dim3(256), 0, st, labels, n_images, hw, rgb);
  return cudaGetLastError();
}

// per-image confusion matrices: conf[b][t][p] for image b, one shared-memory histogram per
// warp; blockIdx.y = image.  Labels outside [0, nc) in EITHER map set *err_flag (the legacy metrics of
// metrics.py:74-183 build their class lists with np.unique over both maps, so every value matters).
__global__ void __launch_bounds__(kHistThreads)
    confusion_batched_kernel(const long long* __restrict__ target, const long long* __restrict__ pred, long long n,
                             int nc, unsigned long long* __restrict__ conf, int* __restrict__ err_flag) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ unsigned int sh[];  // [copies][nc*nc]
  const int nbins = nc * nc;
  const int copies = kHistThreads / 32;
  for (int i = threadIdx.x; i < copies * nbins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  unsigned int* mine = sh + (threadIdx.x >> 5) * nbins;
  const long long* t = target + static_cast<long long>(blockIdx.y) * n;
  const long long* q = pred + static_cast<long long>(blockIdx.y) * n;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long a = t[i], b = q[i];
    if (a >= 0 && a < nc && b >= 0 && b < nc) atomicAdd(mine + a * nc + b, 1u);
    else if (err_flag) *err_flag = 1;
  }
  hist_flush(sh, nbins, copies, conf + static_cast<long long>(blockIdx.y) * nbins);
}
cudaError_t confusion_matrix_batched(const long long* target, const long long* pred, int B, long long n, int nc,
                                     long long* conf, int* err_flag, cudaStream_t st) {
  if (B <= 0 || n <= 0) return cudaSuccess;
  const size_t smem = static_cast<size_t>(kHistThreads / 32) * nc * nc * sizeof(unsigned int);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  long long gx = (n + kHistThreads * 8 - 1) / (kHistThreads * 8);
  if (gx > 64) gx = 64;
  if (gx < 1) gx = 1;
  launch_k(confusion_batched_kernel, dim3(static_cast<unsigned>(gx), B), dim3(kHistThreads), smem, st, target, pred, n, nc,
           reinterpret_cast<unsigned long long*>(conf), err_flag);
  return cudaGetLastError();
}

}  // namespace clk
