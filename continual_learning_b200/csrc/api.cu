// api.cu — the extern "C" boundary of libclk.so (declared in include/clk.h).
// Builds TMA tensor maps + tile geometry for the igemm kernels and forwards the HBM-bound ops.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/clk.h"
#include "igemm.cuh"
#include "membound.cuh"

namespace clk {
int g_pdl = 1;  // programmatic dependent launch for every kernel of the library (clk_ptx.cuh)
}
using namespace clk;

namespace {

thread_local char g_err[512] = "";
int g_fprop_bn = 0;
int g_wgrad_ksplit = 0;
int g_wgrad_ctas = 0;         // generic wgrad: total CTAs aimed at by the split-K heuristic (0 = 2 per SM)
int g_wgrad_bn = 64;
int g_wgrad_v2 = 2;          // conv3x3 wgrad: 0 generic, 1 halo (1 CTA), 2 CTA-pair halo where Cout % 128 == 0 (default)
int g_conv3_v2 = 4;          // conv3x3 fprop/dgrad kernel: 0 generic, 1 hybrid, 2 halo (1 CTA), 4 CTA-pair halo (default)
int g_tma_store = 1;         // generic GEMM kernel: bf16 output through shared-memory staging + TMA tensor stores
int g_conv3_rowtap = 1;      // 64-output-channel conv3x3 forward / dgrad on the row-tap kernel (N = 192), K <= 128
int g_conv3_pair = 1;        // CTA-pair kernel (cta_group::2, BN = 256) whenever the N extent is a multiple of 256
int g_conv3_min_hw = 2048;  // halo kernel for images with at least this many pixels; smaller maps use the generic kernel (BN up to 256)
int g_num_sms_api = 148;
int g_convT_wide = 1;         // ConvTranspose2d forward: 256-column N tiles across quadrants

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int cuda_status(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return CLK_OK;
  return fail(CLK_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
inline cudaStream_t S(clk_stream_t st) { return reinterpret_cast<cudaStream_t>(st); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int ensure_encode() {
  if (g_encode != nullptr) return CLK_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || fn == nullptr || q != cudaDriverEntryPointSuccess)
    return fail(CLK_E_CUDA, "cuTensorMapEncodeTiled entry point unavailable: %s", cudaGetErrorString(e));
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return CLK_OK;
}

// bf16 tensor map, SWIZZLE_128B, zero OOB fill. dims/box innermost first; strides[i] = byte stride of dim i+1.
int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
             const uint32_t* box) {
  int rc = ensure_encode();
  if (rc != CLK_OK) return rc;
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides[i];
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(CLK_E_BADARG, "tensor base not 16-byte aligned");
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), d, s, b, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(CLK_E_CUDA, "cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u,%u]",
                static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0,
                rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0);
  return CLK_OK;
}

int pow2ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// NHWC activation [N][H][W][C] as a 5-D map {C, W, H, N, 1} with box {64, tw, th, nb, 1}
int map_nhwc(CUtensorMap* m, const void* base, int N, int H, int W, int C, int tw, int th, int nb) {
  const uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N, 1};
  const uint64_t st[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2,
                          (uint64_t)N * H * W * C * 2};
  const uint32_t box[5] = {64, (uint32_t)tw, (uint32_t)th, (uint32_t)nb, 1};
  return make_map(m, base, 5, dims, st, box);
}
// [P][C] as {C, P, 1, 1, 1} with box {64, rows, 1, 1, 1}
int map_linear(CUtensorMap* m, const void* base, long long P, int C, int rows) {
  const uint64_t dims[5] = {(uint64_t)C, (uint64_t)P, 1, 1, 1};
  const uint64_t pitch = (uint64_t)P * C * 2;
  const uint64_t st[4] = {(uint64_t)C * 2, pitch, pitch, pitch};
  const uint32_t box[5] = {64, (uint32_t)rows, 1, 1, 1};
  return make_map(m, base, 5, dims, st, box);
}
// stride-2 quadrant view of y [N][2H][2W][C]: {C, 2 (j), W, 2 (i), N*H}, box {64, 1, tw, 1, th}
int map_quad(CUtensorMap* m, const void* base, int N, int H, int W, int C, int tw, int th) {
  const uint64_t dims[5] = {(uint64_t)C, 2, (uint64_t)W, 2, (uint64_t)N * H};
  const uint64_t st[4] = {(uint64_t)C * 2, (uint64_t)2 * C * 2, (uint64_t)2 * W * C * 2,
                          (uint64_t)4 * W * C * 2};
  const uint32_t box[5] = {64, 1, (uint32_t)tw, 1, (uint32_t)th};
  return make_map(m, base, 5, dims, st, box);
}
// rows view [N*H][W][C]: {C, W, N*H, 1, 1}, box {64, tw, th, 1, 1}
int map_rows(CUtensorMap* m, const void* base, int N, int H, int W, int C, int tw, int th) {
  const uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)N * H, 1, 1};
  const uint64_t pitch = (uint64_t)N * H * W * C * 2;
  const uint64_t st[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, pitch, pitch};
  const uint32_t box[5] = {64, (uint32_t)tw, (uint32_t)th, 1, 1};
  return make_map(m, base, 5, dims, st, box);
}
// packed weights [T][Nrows][K] as {K, Nrows, T}, box {64, BN, 1}
int map_weights(CUtensorMap* m, const void* base, int T, int Nrows, int K, int BN) {
  const uint64_t dims[3] = {(uint64_t)K, (uint64_t)Nrows, (uint64_t)T};
  const uint64_t st[2] = {(uint64_t)K * 2, (uint64_t)Nrows * K * 2};
  const uint32_t box[3] = {64, (uint32_t)BN, 1};
  return make_map(m, base, 3, dims, st, box);
}

void geom_nhwc(TileGeom& g, int N, int H, int W, int pixels) {
  memset(&g, 0, sizeof(g));
  g.mode = ADDR_NHWC;
  g.N = N; g.H = H; g.W = W;
  int tw = pow2ceil(W);
  if (tw > 16) tw = 16;
  int th = pow2ceil(H);
  if (th > pixels / tw) th = pixels / tw;
  g.tw = tw; g.th = th; g.nb = pixels / (tw * th);
  g.tiles_w = (W + g.tw - 1) / g.tw;
  g.tiles_h = (H + g.th - 1) / g.th;
  g.tiles_n = (N + g.nb - 1) / g.nb;
}
void geom_linear(TileGeom& g, long long P, int pixels) {
  memset(&g, 0, sizeof(g));
  g.mode = ADDR_LINEAR;
  g.N = 1; g.H = 1; g.W = static_cast<int>(P);
  g.tw = pixels; g.th = 1; g.nb = 1;
  g.tiles_w = static_cast<int>((P + pixels - 1) / pixels);
  g.tiles_h = 1; g.tiles_n = 1;
}
void geom_quad(TileGeom& g, int N, int H, int W, int pixels) {
  memset(&g, 0, sizeof(g));
  g.mode = ADDR_QUAD;
  g.N = N; g.H = H; g.W = W;
  int tw = pow2ceil(W);
  if (tw > 16) tw = 16;
  g.tw = tw; g.th = pixels / tw; g.nb = 1;
  g.tiles_w = (W + g.tw - 1) / g.tw;
  g.tiles_h = (N * H + g.th - 1) / g.th;
  g.tiles_n = 1;
}
void taps_3x3(TileGeom& g) {
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) {
      g.t1[r * 3 + s] = s - 1;
      g.t2[r * 3 + s] = r - 1;
      g.t3[r * 3 + s] = 0;
    }
}
void taps_quad(TileGeom& g) {
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) {
      g.t1[i * 2 + j] = j;
      g.t2[i * 2 + j] = 0;
      g.t3[i * 2 + j] = i;
    }
}
int num_tiles(const TileGeom& g) { return g.tiles_w * g.tiles_h * g.tiles_n; }

int pick_bn(int ncols) {
  if (g_fprop_bn > 0 && ncols % g_fprop_bn == 0) return g_fprop_bn;
  if (ncols % 256 == 0) return 256;
  if (ncols % 128 == 0) return 128;
  return 64;
}
int pick_ksplit(int base_ctas, int tiles_total) {
  if (g_wgrad_ksplit > 0) return g_wgrad_ksplit < tiles_total ? g_wgrad_ksplit : tiles_total;
  int target = g_wgrad_ctas > 0 ? g_wgrad_ctas : 2 * g_num_sms_api;
  int ks = (target + base_ctas - 1) / base_ctas;
  if (ks < 1) ks = 1;
  // keep at least 8 k-steps per CTA so the pipeline fills
  int max_ks = tiles_total / 8;
  if (max_ks < 1) max_ks = 1;
  if (ks > max_ks) ks = max_ks;
  return ks;
}

#define CHECK_RC(x)            \
  do {                         \
    int _rc = (x);             \
    if (_rc != CLK_OK) return _rc; \
  } while (0)

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int clk_version(void) { return 100; }
const char* clk_last_error(void) { return g_err; }

int clk_query_device(int dev) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return cuda_status(e, "cudaGetDeviceProperties");
  if (prop.major != 10)
    return fail(CLK_E_ARCH, "device %d is sm_%d%d; libclk is built for sm_100a only (no fallback)", dev,
                prop.major, prop.minor);
  g_num_sms_api = prop.multiProcessorCount;
  set_num_sms(prop.multiProcessorCount);
  igemm_set_num_sms(prop.multiProcessorCount);
  return CLK_OK;
}

int clk_set_tuning(const char* key, int value) {
  if (key == nullptr) return fail(CLK_E_BADARG, "null key");
  if (strcmp(key, "fprop_bn") == 0) g_fprop_bn = value;
  else if (strcmp(key, "wgrad_ksplit") == 0) g_wgrad_ksplit = value;
  else if (strcmp(key, "wgrad_ctas") == 0) g_wgrad_ctas = value;
  else if (strcmp(key, "wgrad_bn") == 0) g_wgrad_bn = (value == 128 ? 128 : 64);
  else if (strcmp(key, "wgrad_v2") == 0) g_wgrad_v2 = value;
  else if (strcmp(key, "conv3_v2") == 0) g_conv3_v2 = value;
  else if (strcmp(key, "conv3_min_hw") == 0) g_conv3_min_hw = value;
  else if (strcmp(key, "conv3_pair") == 0) g_conv3_pair = value;
  else if (strcmp(key, "conv3_rowtap") == 0) g_conv3_rowtap = value;
  else if (strcmp(key, "tma_store") == 0) g_tma_store = value;
  else if (strcmp(key, "pdl") == 0) g_pdl = value ? 1 : 0;
  else if (strcmp(key, "convT_wide") == 0) g_convT_wide = value ? 1 : 0;
  else if (strcmp(key, "pdl_tensor_trigger") == 0) return cuda_status(igemm_set_pdl_mode(value), "pdl_tensor_trigger");
  else return fail(CLK_E_BADARG, "unknown tuning key %s", key);
  return CLK_OK;
}

// ------------------------------------------------------------------ layout / weights
int clk_nchw_f32_to_nhwc_bf16(const float* x, void* y, int N, int C, int H, int W, int Cpad, clk_stream_t st) {
  if (!x || !y || N <= 0 || C <= 0 || Cpad < C) return fail(CLK_E_BADARG, "nchw_f32_to_nhwc_bf16: bad args");
  return cuda_status(nchw_f32_to_nhwc_bf16(x, y, N, C, H, W, Cpad, S(st)), "nchw_f32_to_nhwc_bf16");
}
int clk_nhwc_to_nchw_f32(const void* x, int x_is_f32, float* y, int N, int C, int H, int W, int ldc,
                         clk_stream_t st) {
  if (!x || !y || N <= 0 || C <= 0 || ldc < C) return fail(CLK_E_BADARG, "nhwc_to_nchw_f32: bad args");
  return cuda_status(nhwc_to_nchw_f32(x, x_is_f32, y, N, C, H, W, ldc, S(st)), "nhwc_to_nchw_f32");
}
int clk_im2col3x3_stem(const float* x, void* a, int N, int Cin, int H, int W, clk_stream_t st) {
  if (!x || !a || N <= 0 || H <= 0 || W <= 0) return fail(CLK_E_BADARG, "im2col3x3_stem: bad args");
  if (Cin * 9 > 64) return fail(CLK_E_UNSUPPORTED_SHAPE, "im2col3x3_stem: Cin*9 must be <= 64 (Cin=%d)", Cin);
  return cuda_status(im2col3x3_stem(x, a, N, Cin, H, W, S(st)), "im2col3x3_stem");
}
int clk_pack_w(const float* src, void* outAB, void* outBA, int A, int B, int T, int ldA, int ldB, int ldB2,
               int ldA2, int rev, clk_stream_t st) {
  if (!src || A <= 0 || B <= 0 || T <= 0 || T > 9) return fail(CLK_E_BADARG, "pack_w: bad args");
  return cuda_status(pack_w(src, outAB, outBA, A, B, T, ldA, ldB, ldB2, ldA2, rev, S(st)), "pack_w");
}
int clk_unpack_wgrad(const float* D, float* grad, int A, int B, int T, int ldA, int ldB, float alpha,
                     int accumulate, int transposed, clk_stream_t st) {
  if (!D || !grad || A <= 0 || B <= 0 || T <= 0 || T > 9) return fail(CLK_E_BADARG, "unpack_wgrad: bad args");
  return cuda_status(unpack_wgrad(D, grad, A, B, T, ldA, ldB, alpha, accumulate, transposed, S(st)), "unpack_wgrad");
}

int clk_pack_w_multi(const void* jobs, int njobs, int total_tiles, int max_T, clk_stream_t st) {
  if (!jobs || njobs <= 0 || total_tiles < 0 || max_T <= 0 || max_T > 9) return fail(CLK_E_BADARG, "pack_w_multi: bad args");
  return cuda_status(pack_w_multi(jobs, njobs, total_tiles, max_T, S(st)), "pack_w_multi");
}
int clk_unpack_wgrad_multi(const void* jobs, int njobs, int total_tiles, int max_T, clk_stream_t st) {
  if (!jobs || njobs <= 0 || total_tiles < 0 || max_T <= 0 || max_T > 9) return fail(CLK_E_BADARG, "unpack_wgrad_multi: bad args");
  return cuda_status(unpack_wgrad_multi(jobs, njobs, total_tiles, max_T, S(st)), "unpack_wgrad_multi");
}
int clk_reduce_partials_multi(const void* jobs, int njobs, int total_blocks, clk_stream_t st) {
  if (!jobs || njobs <= 0 || total_blocks < 0) return fail(CLK_E_BADARG, "reduce_partials_multi: bad args");
  return cuda_status(reduce_partials_multi(jobs, njobs, total_blocks, S(st)), "reduce_partials_multi");
}
int clk_f64_to_f32_multi(const void* jobs, int njobs, clk_stream_t st) {
  if (!jobs || njobs <= 0) return fail(CLK_E_BADARG, "f64_to_f32_multi: bad args");
  return cuda_status(f64_to_f32_multi(jobs, njobs, S(st)), "f64_to_f32_multi");
}

// ------------------------------------------------------------------ igemm: conv3x3
static int conv3x3_fprop_impl(const void* x0, int C0, const void* x1, int C1, const void* w, const float* bias,
                             void* y, double* stat_sum, double* stat_sq, const float* bn_scale,
                             const float* bn_shift, int N, int H, int W, int Cout, int relu, clk_stream_t st) {
  if (!x0 || !w || !y || N <= 0 || H <= 0 || W <= 0) return fail(CLK_E_BADARG, "conv3x3_fprop: bad args");
  if ((bn_scale == nullptr) != (bn_shift == nullptr) || (bn_scale && stat_sum))
    return fail(CLK_E_BADARG, "conv3x3_fprop: scale and shift come together and exclude the statistics");
  if (C0 % 64 || C1 % 64 || Cout % 64 || C0 <= 0 || (x1 == nullptr) != (C1 == 0))
    return fail(CLK_E_UNSUPPORTED_SHAPE, "conv3x3_fprop: channels must be multiples of 64 (C0=%d C1=%d Cout=%d)",
                C0, C1, Cout);
  const bool pair_f = g_conv3_v2 == 4 || (g_conv3_v2 && g_conv3_pair && Cout % 256 == 0 && g_fprop_bn == 0);
  if (pair_f || (g_conv3_v2 && (H * W >= g_conv3_min_hw || g_conv3_v2 == 2))) {
    Conv3Params q;
    memset(&q, 0, sizeof(q));
    q.N = N; q.H = H; q.W = W;
    q.tiles_w = (W + 15) / 16;
    q.tiles_h = (H + 15) / 16;
    q.m_tiles = N * q.tiles_h * q.tiles_w;
    const bool pair = pair_f;  // CTA-pair kernel (cta_group::2)
    int BNq = (Cout % 128 == 0 && g_fprop_bn != 64) ? 128 : 64;
    if (pair && Cout % 256 == 0 && g_fprop_bn != 64 && g_fprop_bn != 128) BNq = 256;
    q.n_tiles = Cout / BNq;
    q.kc0 = C0 / 64;
    q.kc1 = C1 / 64;
    q.n_store = Cout;
    q.dst0 = y;
    q.ldc0 = Cout;
    q.bias = bias;
    q.relu = relu;
    q.stat_sum = stat_sum;
    q.stat_sq = stat_sq;
    q.bn_scale = bn_scale;
    q.bn_shift = bn_shift;
    CUtensorMap a0, a1, b;
    if (pair && g_conv3_rowtap && Cout == 64 && q.kc0 + q.kc1 <= 2) {
      // row-tap kernel: the three horizontal taps of a kernel row in one N = 192 MMA (igemm_conv3r_kernel)
      q.tiles_w = (W + 29) / 30;
      q.tiles_h = (H + 7) / 8;
      q.m_tiles = N * q.tiles_h * q.tiles_w;
      q.n_tiles = 1;
      CHECK_RC(map_nhwc(&a0, x0, N, H, W, C0, 32, 6, 1));
      if (x1) CHECK_RC(map_nhwc(&a1, x1, N, H, W, C1, 32, 6, 1));
      else a1 = a0;
      CHECK_RC(map_weights(&b, w, 3, 192, C0 + C1, 32));
      CUtensorMap o;
      CHECK_RC(map_nhwc(&o, y, N, H, W, 64, 30, 4, 1));
      return cuda_status(launch_conv3r(a0, a1, b, o, q, g_num_sms_api, S(st)), "conv3x3_fprop(row-tap)");
    }
    int sub = BNq == 256 ? 1 : 2;
    if (pair && BNq == 256 && 2 * q.m_tiles * q.n_tiles <= g_num_sms_api / 2) {
      BNq = 128;  // fewer tiles than half the clusters: split N finer (one 16x16-pixel sub tile, 128 columns)
      q.n_tiles = Cout / BNq;
    }
    const int bw = (pair && sub == 1) ? 16 : 24;
    if (pair && sub == 2) {  // two sub tiles per CTA: the pair covers 16 x 32 pixels
      q.tiles_w = (W + 31) / 32;
      q.m_tiles = N * q.tiles_h * q.tiles_w;
    }
    CHECK_RC(map_nhwc(&a0, x0, N, H, W, C0, bw, 18, 1));
    if (x1) CHECK_RC(map_nhwc(&a1, x1, N, H, W, C1, bw, 18, 1));
    else a1 = a0;
    CHECK_RC(map_weights(&b, w, 9, Cout, C0 + C1, pair ? BNq / 2 : BNq));
    if (pair) return cuda_status(launch_conv3x2(BNq, sub, a0, a1, b, q, g_num_sms_api, S(st)), "conv3x3_fprop(pair)");
    return cuda_status(launch_conv3(BNq, a0, a1, b, q, g_num_sms_api, S(st)), "conv3x3_fprop(halo)");
  }
  FpropParams p;
  memset(&p, 0, sizeof(p));
  geom_nhwc(p.g, N, H, W, 128);
  taps_3x3(p.g);
  p.ntaps = 9;
  p.kc0 = C0 / 64;
  p.kc1 = C1 / 64;
  p.n_store = Cout;
  p.dst0 = y;
  p.ldc0 = Cout;
  p.bias = bias;
  p.relu = relu;
  p.stat_sum = stat_sum;
  p.stat_sq = stat_sq;
  p.bn_scale = bn_scale;
  p.bn_shift = bn_shift;
  const int BN = pick_bn(Cout);
  CUtensorMap a0, a1, b;
  CHECK_RC(map_nhwc(&a0, x0, N, H, W, C0, p.g.tw, p.g.th, p.g.nb));
  if (x1) CHECK_RC(map_nhwc(&a1, x1, N, H, W, C1, p.g.tw, p.g.th, p.g.nb));
  else a1 = a0;
  CHECK_RC(map_weights(&b, w, 9, Cout, C0 + C1, BN));
  return cuda_status(launch_fprop(BN, 0, a0, a1, b, p, num_tiles(p.g), Cout / BN, S(st)), "conv3x3_fprop");
}

int clk_conv3x3_fprop(const void* x0, int C0, const void* x1, int C1, const void* w, const float* bias,
                      void* y, double* stat_sum, double* stat_sq, int N, int H, int W, int Cout, int relu,
                      clk_stream_t st) {
  return conv3x3_fprop_impl(x0, C0, x1, C1, w, bias, y, stat_sum, stat_sq, nullptr, nullptr, N, H, W, Cout, relu, st);
}

int clk_conv3x3_fprop_eval(const void* x0, int C0, const void* x1, int C1, const void* w, const float* bias,
                           void* z, const float* bn_scale, const float* bn_shift, int N, int H, int W, int Cout,
                           int relu, clk_stream_t st) {
  if (!bn_scale || !bn_shift) return fail(CLK_E_BADARG, "conv3x3_fprop_eval: scale / shift missing");
  return conv3x3_fprop_impl(x0, C0, x1, C1, w, bias, z, nullptr, nullptr, bn_scale, bn_shift, N, H, W, Cout, relu, st);
}

int clk_conv3x3_dgrad(const void* dy, int Cout, const void* wd, void* dx0, int C0, void* dx1, int C1, int N,
                      int H, int W, clk_stream_t st) {
  if (!dy || !wd || !dx0 || N <= 0 || H <= 0 || W <= 0) return fail(CLK_E_BADARG, "conv3x3_dgrad: bad args");
  if (C0 % 64 || C1 % 64 || Cout % 64 || C0 <= 0 || (dx1 == nullptr) != (C1 == 0))
    return fail(CLK_E_UNSUPPORTED_SHAPE, "conv3x3_dgrad: channels must be multiples of 64");
  const bool pair_d = g_conv3_v2 == 4 || (g_conv3_v2 && g_conv3_pair && (C0 + C1) % 256 == 0 && g_fprop_bn == 0);
  if (pair_d || (g_conv3_v2 && (H * W >= g_conv3_min_hw || g_conv3_v2 == 2))) {
    Conv3Params q;
    memset(&q, 0, sizeof(q));
    const int Cin2 = C0 + C1;
    q.N = N; q.H = H; q.W = W;
    q.tiles_w = (W + 15) / 16;
    q.tiles_h = (H + 15) / 16;
    q.m_tiles = N * q.tiles_h * q.tiles_w;
    const bool pair = pair_d;
    int BNq = (Cin2 % 128 == 0 && g_fprop_bn != 64) ? 128 : 64;  // a tile may straddle the C0 | C1 split
    if (pair && Cin2 % 256 == 0 && g_fprop_bn != 64 && g_fprop_bn != 128) BNq = 256;
    q.n_tiles = Cin2 / BNq;
    q.kc0 = Cout / 64;
    q.kc1 = 0;
    q.n_store = Cin2;
    q.split_c = C1 > 0 ? C0 : 0;
    q.dst0 = dx0;
    q.ldc0 = C0;
    q.dst1 = dx1;
    q.ldc1 = C1;
    CUtensorMap a0, b;
    if (pair && g_conv3_rowtap && Cin2 == 64 && q.kc0 <= 2) {
      q.tiles_w = (W + 29) / 30;
      q.tiles_h = (H + 7) / 8;
      q.m_tiles = N * q.tiles_h * q.tiles_w;
      q.n_tiles = 1;
      CHECK_RC(map_nhwc(&a0, dy, N, H, W, Cout, 32, 6, 1));
      CHECK_RC(map_weights(&b, wd, 3, 192, Cout, 32));
      CUtensorMap o;
      CHECK_RC(map_nhwc(&o, dx0, N, H, W, 64, 30, 4, 1));
      return cuda_status(launch_conv3r(a0, a0, b, o, q, g_num_sms_api, S(st)), "conv3x3_dgrad(row-tap)");
    }
    int sub = BNq == 256 ? 1 : 2;
    if (pair && BNq == 256 && 2 * q.m_tiles * q.n_tiles <= g_num_sms_api / 2) {
      BNq = 128;  // fewer tiles than half the clusters: split N finer (one 16x16-pixel sub tile, 128 columns)
      q.n_tiles = Cin2 / BNq;
    }
    if (pair && sub == 2) {
      q.tiles_w = (W + 31) / 32;
      q.m_tiles = N * q.tiles_h * q.tiles_w;
    }
    CHECK_RC(map_nhwc(&a0, dy, N, H, W, Cout, (pair && sub == 1) ? 16 : 24, 18, 1));
    CHECK_RC(map_weights(&b, wd, 9, Cin2, Cout, pair ? BNq / 2 : BNq));
    if (pair) return cuda_status(launch_conv3x2(BNq, sub, a0, a0, b, q, g_num_sms_api, S(st)), "conv3x3_dgrad(pair)");
    return cuda_status(launch_conv3(BNq, a0, a0, b, q, g_num_sms_api, S(st)), "conv3x3_dgrad(halo)");
  }
  FpropParams p;
  memset(&p, 0, sizeof(p));
  geom_nhwc(p.g, N, H, W, 128);
  taps_3x3(p.g);
  p.ntaps = 9;
  p.kc0 = Cout / 64;
  p.kc1 = 0;
  const int Cin = C0 + C1;
  p.n_store = Cin;
  p.dst0 = dx0;
  p.ldc0 = C0;
  p.dst1 = dx1;
  p.ldc1 = C1;
  p.split_c = C1 > 0 ? C0 : 0;
  int BN = pick_bn(Cin);
  if (C1 > 0) {
    while (BN > 64 && (C0 % BN != 0 || C1 % BN != 0)) BN >>= 1;
  }
  CUtensorMap a0, b;
  CHECK_RC(map_nhwc(&a0, dy, N, H, W, Cout, p.g.tw, p.g.th, p.g.nb));
  CHECK_RC(map_weights(&b, wd, 9, Cin, Cout, BN));
  return cuda_status(launch_fprop(BN, 0, a0, a0, b, p, num_tiles(p.g), Cin / BN, S(st)), "conv3x3_dgrad");
}

static int wgrad9_plan(int Cout, int Cin, int N, int H, int W, Wgrad9Params& q, bool& pair) {
  memset(&q, 0, sizeof(q));
  q.N = N; q.H = H; q.W = W;
  q.tiles_w = (W + 7) / 8;
  q.tiles_h = (H + 15) / 16;
  q.tiles_total = N * q.tiles_h * q.tiles_w;
  q.cin_slabs = Cin / 64;
  q.cout_tiles = Cout / 64;
  q.Cin = Cin;
  q.Cout = Cout;
  pair = g_wgrad_v2 == 2 && Cout % 128 == 0;  // CTA-pair kernel (cta_group::2)
  if (pair) q.cout_tiles = Cout / 128;
  const int base = q.cin_slabs * q.cout_tiles;
  int ks = g_wgrad_ksplit > 0 ? g_wgrad_ksplit : ((pair ? g_num_sms_api / 2 : g_num_sms_api) / base);
  if (ks < 1) ks = 1;
  if (ks > q.tiles_total) ks = q.tiles_total;
  const int per = (q.tiles_total + ks - 1) / ks;
  q.ksplit = (q.tiles_total + per - 1) / per;  // every split owns at least one pixel tile
  return q.ksplit;
}

int clk_conv3x3_wgrad_splits(int Cout, int Cin, int N, int H, int W) {
  if (Cout % 64 || Cin % 64 || Cout <= 0 || Cin <= 0 || N <= 0 || H <= 0 || W <= 0)
    return fail(CLK_E_UNSUPPORTED_SHAPE, "conv3x3_wgrad_splits: channels must be multiples of 64");
  if (!g_wgrad_v2) return 1;
  Wgrad9Params q;
  bool pair;
  return wgrad9_plan(Cout, Cin, N, H, W, q, pair);
}

int clk_conv3x3_wgrad_split(const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1, float* partials,
                            int N, int H, int W, clk_stream_t st) {
  if (!dy || !x0 || !partials || N <= 0 || H <= 0 || W <= 0) return fail(CLK_E_BADARG, "conv3x3_wgrad_split: bad args");
  if (C0 % 64 || C1 % 64 || Cout % 64 || C0 <= 0 || (x1 == nullptr) != (C1 == 0))
    return fail(CLK_E_UNSUPPORTED_SHAPE, "conv3x3_wgrad_split: channels must be multiples of 64");
  if (!g_wgrad_v2) return fail(CLK_E_UNSUPPORTED_SHAPE, "conv3x3_wgrad_split needs the halo kernels (wgrad_v2 != 0)");
  Wgrad9Params q;
  bool pair;
  wgrad9_plan(Cout, C0 + C1, N, H, W, q, pair);
  q.split_slabs = C0 / 64;
  q.out = partials;
  q.split_stride = 9LL * (C0 + C1) * Cout;
  CUtensorMap u, t0, t1;
  CHECK_RC(map_nhwc(&u, dy, N, H, W, Cout, 8, 16, 1));
  CHECK_RC(map_nhwc(&t0, x0, N, H, W, C0, 16, 18, 1));
  if (x1) CHECK_RC(map_nhwc(&t1, x1, N, H, W, C1, 16, 18, 1));
  else t1 = t0;
  if (pair) return cuda_status(launch_wgrad9x2(u, t0, t1, q, S(st)), "conv3x3_wgrad_split(pair)");
  return cuda_status(launch_wgrad9(u, t0, t1, q, S(st)), "conv3x3_wgrad_split(halo)");
}

int clk_conv3x3_wgrad(const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1, float* dw,
                      int N, int H, int W, clk_stream_t st) {
  if (!dy || !x0 || !dw || N <= 0 || H <= 0 || W <= 0) return fail(CLK_E_BADARG, "conv3x3_wgrad: bad args");
  if (C0 % 64 || C1 % 64 || Cout % 64 || C0 <= 0 || (x1 == nullptr) != (C1 == 0))
    return fail(CLK_E_UNSUPPORTED_SHAPE, "conv3x3_wgrad: channels must be multiples of 64");
  if (g_wgrad_v2) {
    // halo variants: all nine taps per CTA (or CTA pair), 16x8-pixel K tiles, split-K accumulated with vector REDs
    Wgrad9Params q;
    bool pair;
    wgrad9_plan(Cout, C0 + C1, N, H, W, q, pair);
    q.split_slabs = C0 / 64;
    q.out = dw;
    q.split_stride = 0;
    CUtensorMap u, t0, t1;
    CHECK_RC(map_nhwc(&u, dy, N, H, W, Cout, 8, 16, 1));
    CHECK_RC(map_nhwc(&t0, x0, N, H, W, C0, 16, 18, 1));
    if (x1) CHECK_RC(map_nhwc(&t1, x1, N, H, W, C1, 16, 18, 1));
    else t1 = t0;
    if (pair) return cuda_status(launch_wgrad9x2(u, t0, t1, q, S(st)), "conv3x3_wgrad(pair)");
    return cuda_status(launch_wgrad9(u, t0, t1, q, S(st)), "conv3x3_wgrad(halo)");
  }
  WgradParams p;
  memset(&p, 0, sizeof(p));
  geom_nhwc(p.g, N, H, W, 64);
  taps_3x3(p.g);
  const int Cin = C0 + C1;
  int BN = g_wgrad_bn;
  if (Cin % BN || (C1 > 0 && C0 % BN)) BN = 64;
  p.ntaps = 9;
  p.G = 3;
  p.CU = Cout;
  p.CT = Cin;
  p.ct_split = C1 > 0 ? C0 : Cin;
  p.m_tiles = (Cout + 127) / 128;
  p.n_tiles = Cin / BN;
  p.tap_groups = 3;
  p.tiles_total = num_tiles(p.g);
  p.ksplit = pick_ksplit(p.m_tiles * p.n_tiles * p.tap_groups, p.tiles_total);
  p.out = dw;
  p.ld_u = Cout;
  p.ld_t = Cin;
  p.transpose_out = 1;  // same packed layout as the halo kernel: [9][Cin][Cout]
  CUtensorMap u, t0, t1;
  CHECK_RC(map_nhwc(&u, dy, N, H, W, Cout, p.g.tw, p.g.th, p.g.nb));
  CHECK_RC(map_nhwc(&t0, x0, N, H, W, C0, p.g.tw, p.g.th, p.g.nb));
  if (x1) CHECK_RC(map_nhwc(&t1, x1, N, H, W, C1, p.g.tw, p.g.th, p.g.nb));
  else t1 = t0;
  return cuda_status(launch_wgrad(BN, u, t0, t1, p, S(st)), "conv3x3_wgrad");
}

// ------------------------------------------------------------------ fused head + loss + head backward
int clk_head_loss_bwd(const void* z, const void* wf, const void* wd, const float* bias, const int64_t* labels,
                      const float* old_logits, long long P, int Cin, int C, int Cold, float T, float lambda,
                      float gscale, void* dz, float* dw, double* dbias, double* loss_acc, int* err_flag,
                      clk_stream_t st) {
  if (!z || !wf || !wd || !labels || !dz || !dw || !dbias || !loss_acc || P <= 0)
    return fail(CLK_E_BADARG, "head_loss_bwd: bad args");
  if (Cin != 64 || C < 1 || C > 32 || Cold < 0 || Cold > C || (old_logits != nullptr) != (Cold > 0))
    return fail(CLK_E_UNSUPPORTED_SHAPE, "head_loss_bwd: needs Cin == 64, 1 <= C <= 32, Cold <= C (Cin=%d C=%d Cold=%d)",
                Cin, C, Cold);
  if (P > 2000000000LL) return fail(CLK_E_UNSUPPORTED_SHAPE, "head_loss_bwd: P too large");
  if (old_logits != nullptr && !(T > 0.f)) return fail(CLK_E_BADARG, "head_loss_bwd: T must be positive");
  HeadLossParams q;
  memset(&q, 0, sizeof(q));
  q.P = P; q.C = C; q.Cold = Cold; q.T = old_logits ? T : 1.f; q.lambda = lambda; q.gscale = gscale;
  q.bias = bias; q.labels = reinterpret_cast<const long long*>(labels); q.old_logits = old_logits;
  q.dz = dz; q.dw = dw; q.dbias = dbias; q.loss_acc = loss_acc; q.err_flag = err_flag;
  CUtensorMap mz, mf, md;
  CHECK_RC(map_linear(&mz, z, P, 64, 128));
  CHECK_RC(map_weights(&mf, wf, 1, 32, 64, 32));
  CHECK_RC(map_weights(&md, wd, 1, 64, 64, 64));
  return cuda_status(launch_head_loss(mz, mf, md, q, g_num_sms_api, S(st)), "head_loss_bwd");
}

// ------------------------------------------------------------------ data contract around the step
int clk_confusion_matrix_batched(const int64_t* target, const int64_t* pred, int B, long long n, int nc, int64_t* conf,
                                 int* err_flag, clk_stream_t st) {
  if (!target || !pred || !conf || B < 0 || n < 0 || nc < 1) return fail(CLK_E_BADARG, "confusion_matrix_batched: bad args");
  if (nc > 38) return fail(CLK_E_UNSUPPORTED_SHAPE, "confusion_matrix_batched: nc <= 38 (nc=%d)", nc);
  return cuda_status(confusion_matrix_batched(reinterpret_cast<const long long*>(target),
                                              reinterpret_cast<const long long*>(pred), B, n, nc,
                                              reinterpret_cast<long long*>(conf), err_flag, S(st)),
                     "confusion_matrix_batched");
}

int clk_voc_prepare_batch(const void* items, int B, int H, int W, float* x, int64_t* y, int* err_flag,
                          clk_stream_t st) {
  if (!items || B < 0 || H <= 0 || W <= 0 || (!x && !y)) return fail(CLK_E_BADARG, "voc_prepare_batch: bad args");
  return cuda_status(voc_prepare_batch(items, B, H, W, x, reinterpret_cast<long long*>(y), err_flag, S(st)),
                     "voc_prepare_batch");
}

int clk_labels_to_rgb(const int64_t* labels, long long n_images, long long hw, double* rgb, clk_stream_t st) {
  if (!labels || !rgb || n_images < 0 || hw < 0) return fail(CLK_E_BADARG, "labels_to_rgb: bad args");
  return cuda_status(labels_to_rgb(reinterpret_cast<const long long*>(labels), n_images, hw, rgb, S(st)),
                     "labels_to_rgb");
}

int clk_head_argmax_confusion(const void* z, const void* wf, const float* bias, const int64_t* labels, long long P,
                              int Cin, int C, int nc, int64_t* pred_out, int64_t* conf, int64_t* correct,
                              clk_stream_t st) {
  if (!z || !wf || !labels || P <= 0) return fail(CLK_E_BADARG, "head_argmax_confusion: bad args");
  if (Cin != 64 || C < 1 || C > 32 || nc < C || nc > 64)
    return fail(CLK_E_UNSUPPORTED_SHAPE, "head_argmax_confusion: needs Cin == 64, 1 <= C <= 32, C <= nc <= 64 (Cin=%d C=%d nc=%d)",
                Cin, C, nc);
  if (P > 2000000000LL) return fail(CLK_E_UNSUPPORTED_SHAPE, "head_argmax_confusion: P too large");
  HeadArgmaxParams q;
  memset(&q, 0, sizeof(q));
  q.P = P; q.C = C; q.nc = nc; q.bias = bias; q.labels = reinterpret_cast<const long long*>(labels);
  q.pred_out = reinterpret_cast<long long*>(pred_out);
  q.conf = reinterpret_cast<unsigned long long*>(conf);
  q.correct = reinterpret_cast<unsigned long long*>(correct);
  CUtensorMap mz, mf;
  CHECK_RC(map_linear(&mz, z, P, 64, 128));
  CHECK_RC(map_weights(&mf, wf, 1, 32, 64, 32));
  return cuda_status(launch_head_argmax(mz, mf, q, g_num_sms_api, S(st)), "head_argmax_confusion");
}

// ------------------------------------------------------------------ igemm: plain GEMMs
static int gemm_fprop_impl(const void* a, int K, const void* w, const float* bias, void* out, int ldo, int n_store,
                          int out_is_f32, int relu, double* stat_sum, double* stat_sq, const float* bn_scale,
                          const float* bn_shift, long long P, int Npad, clk_stream_t st) {
  if (!a || !w || !out || P <= 0) return fail(CLK_E_BADARG, "gemm_fprop: bad args");
  if ((bn_scale == nullptr) != (bn_shift == nullptr) || (bn_scale && stat_sum))
    return fail(CLK_E_BADARG, "gemm_fprop: scale and shift come together and exclude the statistics");
  if (K % 64 || K <= 0) return fail(CLK_E_UNSUPPORTED_SHAPE, "gemm_fprop: K must be a multiple of 64 (K=%d)", K);
  if (P > 2000000000LL) return fail(CLK_E_UNSUPPORTED_SHAPE, "gemm_fprop: P too large");
  int BN;
  if (out_is_f32) {
    if (Npad != 32) return fail(CLK_E_UNSUPPORTED_SHAPE, "gemm_fprop: fp32 output needs Npad == 32");
    BN = 32;
  } else {
    if (Npad % 64) return fail(CLK_E_UNSUPPORTED_SHAPE, "gemm_fprop: Npad must be a multiple of 64");
    BN = pick_bn(Npad);
  }
  FpropParams p;
  memset(&p, 0, sizeof(p));
  geom_linear(p.g, P, 128);
  p.ntaps = 1;
  p.kc0 = K / 64;
  p.n_store = n_store;
  p.dst0 = out;
  p.ldc0 = ldo;
  p.bias = bias;
  p.relu = relu;
  p.stat_sum = stat_sum;
  p.stat_sq = stat_sq;
  p.bn_scale = bn_scale;
  p.bn_shift = bn_shift;
  CUtensorMap a0, b;
  CHECK_RC(map_linear(&a0, a, P, K, 128));
  CHECK_RC(map_weights(&b, w, 1, Npad, K, BN));
  if (g_tma_store && !out_is_f32 && n_store % 64 == 0 && ldo % 8 == 0) {
    CUtensorMap o;
    CHECK_RC(map_linear(&o, out, P, ldo, 128));
    p.tma_store = 1;
    return cuda_status(launch_fprop(BN, 0, a0, a0, b, p, num_tiles(p.g), Npad / BN, S(st), &o), "gemm_fprop");
  }
  return cuda_status(launch_fprop(BN, out_is_f32, a0, a0, b, p, num_tiles(p.g), Npad / BN, S(st)), "gemm_fprop");
}

int clk_stem_conv3x3_fprop(const float* x, int Cin, const void* w, const float* bias, void* y, double* stat_sum,
                           double* stat_sq, const float* bn_scale, const float* bn_shift, int N, int H, int W, int relu,
                           clk_stream_t st) {
  if (!x || !w || !y || N <= 0 || H <= 0 || W <= 0 || Cin <= 0) return fail(CLK_E_BADARG, "stem_conv3x3_fprop: bad args");
  if ((bn_scale == nullptr) != (bn_shift == nullptr) || (bn_scale && stat_sum) || (stat_sum == nullptr) != (stat_sq == nullptr))
    return fail(CLK_E_BADARG, "stem_conv3x3_fprop: scale and shift come together and exclude the statistics");
  if (Cin * 9 > 32 || H % 8 || W % 16)
    return fail(CLK_E_UNSUPPORTED_SHAPE, "stem_conv3x3_fprop: needs Cin*9 <= 32, H %% 8 == 0, W %% 16 == 0 (Cin=%d H=%d W=%d); "
                "use clk_im2col3x3_stem + clk_gemm_fprop", Cin, H, W);
  CUtensorMap mw, mo;
  CHECK_RC(map_weights(&mw, w, 1, 64, 64, 64));
  CHECK_RC(map_nhwc(&mo, y, N, H, W, 64, 16, 8, 1));
  return cuda_status(launch_stem_conv(mw, mo, x, N, Cin, H, W, bias, relu, stat_sum, stat_sq, bn_scale, bn_shift,
                                      g_num_sms_api, S(st)), "stem_conv3x3_fprop");
}

int clk_gemm_fprop(const void* a, int K, const void* w, const float* bias, void* out, int ldo, int n_store,
                   int out_is_f32, int relu, double* stat_sum, double* stat_sq, long long P, int Npad,
                   clk_stream_t st) {
  return gemm_fprop_impl(a, K, w, bias, out, ldo, n_store, out_is_f32, relu, stat_sum, stat_sq, nullptr, nullptr, P,
                         Npad, st);
}

int clk_gemm_fprop_eval(const void* a, int K, const void* w, const float* bias, void* out, int ldo, int n_store,
                        int relu, const float* bn_scale, const float* bn_shift, long long P, int Npad,
                        clk_stream_t st) {
  if (!bn_scale || !bn_shift) return fail(CLK_E_BADARG, "gemm_fprop_eval: scale / shift missing");
  return gemm_fprop_impl(a, K, w, bias, out, ldo, n_store, 0, relu, nullptr, nullptr, bn_scale, bn_shift, P, Npad, st);
}

int clk_gemm_wgrad(const void* u, int CU, const void* t, int CT, float* out, int ld_u, int ld_t, long long P,
                   clk_stream_t st) {
  if (!u || !t || !out || P <= 0) return fail(CLK_E_BADARG, "gemm_wgrad: bad args");
  if (CU % 64 || CT % 64) return fail(CLK_E_UNSUPPORTED_SHAPE, "gemm_wgrad: channels must be multiples of 64");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  geom_linear(p.g, P, 64);
  p.ntaps = 1;
  p.G = 1;
  p.CU = CU;
  p.CT = CT;
  p.ct_split = CT;
  p.m_tiles = (CU + 127) / 128;
  p.n_tiles = CT / 64;
  p.tap_groups = 1;
  p.tiles_total = num_tiles(p.g);
  p.ksplit = pick_ksplit(p.m_tiles * p.n_tiles, p.tiles_total);
  p.out = out;
  p.ld_u = ld_u;
  p.ld_t = ld_t;
  CUtensorMap mu, mt;
  CHECK_RC(map_linear(&mu, u, P, CU, 64));
  CHECK_RC(map_linear(&mt, t, P, CT, 64));
  return cuda_status(launch_wgrad(64, mu, mt, mt, p, S(st)), "gemm_wgrad");
}

// ------------------------------------------------------------------ igemm: ConvTranspose2d 2x2 s2
int clk_convT2x2_fprop(const void* x, const void* w, const float* bias, void* y, int N, int H, int W, int Cin,
                       int Cout, clk_stream_t st) {
  if (!x || !w || !y || N <= 0 || H <= 0 || W <= 0) return fail(CLK_E_BADARG, "convT2x2_fprop: bad args");
  if (Cin % 64 || Cout % 64) return fail(CLK_E_UNSUPPORTED_SHAPE, "convT2x2_fprop: channels must be multiples of 64");
  const long long P = static_cast<long long>(N) * H * W;
  FpropParams p;
  memset(&p, 0, sizeof(p));
  geom_linear(p.g, P, 128);
  p.g.N = N; p.g.H = H; p.g.W = W;  // pixel decode for the shuffle epilogue
  p.ntaps = 1;
  p.kc0 = Cin / 64;
  p.shuffle = 1;
  p.cout_q = Cout;
  p.n_store = 4 * Cout;
  p.dst0 = y;
  p.ldc0 = Cout;
  p.bias = bias;
  // N tiles of 256 columns span several quadrants when Cout < 256: the input tile is read once per 256 output
  // columns instead of once per quadrant (the epilogue picks the destination pixel per 32-column chunk)
  const int BN = g_convT_wide ? 256 : pick_bn(Cout);
  CUtensorMap a0, b;
  CHECK_RC(map_linear(&a0, x, P, Cin, 128));
  CHECK_RC(map_weights(&b, w, 1, 4 * Cout, Cin, BN));
  if (g_tma_store && (W % 128 == 0 || 128 % W == 0)) {
    // a tile = 128 consecutive input pixels = a [th][tw] block of the merged [N*H][W] grid: one TMA store per
    // 64-column group through the quadrant view {C, 2, W, 2, N*H} of y
    const int tw = W < 128 ? W : 128;
    CUtensorMap o;
    CHECK_RC(map_quad(&o, y, N, H, W, Cout, tw, 128 / tw));
    p.tma_store = 2;
    return cuda_status(launch_fprop(BN, 0, a0, a0, b, p, num_tiles(p.g), 4 * Cout / BN, S(st), &o), "convT2x2_fprop");
  }
  return cuda_status(launch_fprop(BN, 0, a0, a0, b, p, num_tiles(p.g), 4 * Cout / BN, S(st)), "convT2x2_fprop");
}

int clk_convT2x2_dgrad(const void* dy, const void* wd, void* dx, int N, int H, int W, int Cin, int Cout,
                       clk_stream_t st) {
  if (!dy || !wd || !dx || N <= 0 || H <= 0 || W <= 0) return fail(CLK_E_BADARG, "convT2x2_dgrad: bad args");
  if (Cin % 64 || Cout % 64) return fail(CLK_E_UNSUPPORTED_SHAPE, "convT2x2_dgrad: channels must be multiples of 64");
  FpropParams p;
  memset(&p, 0, sizeof(p));
  geom_quad(p.g, N, H, W, 128);
  taps_quad(p.g);
  p.ntaps = 4;
  p.kc0 = Cout / 64;
  p.n_store = Cin;
  p.dst0 = dx;
  p.ldc0 = Cin;
  const int BN = pick_bn(Cin);
  CUtensorMap a0, b;
  CHECK_RC(map_quad(&a0, dy, N, H, W, Cout, p.g.tw, p.g.th));
  CHECK_RC(map_weights(&b, wd, 4, Cin, Cout, BN));
  if (g_tma_store) {
    CUtensorMap o;
    CHECK_RC(map_rows(&o, dx, N, H, W, Cin, p.g.tw, p.g.th));
    p.tma_store = 3;
    return cuda_status(launch_fprop(BN, 0, a0, a0, b, p, num_tiles(p.g), Cin / BN, S(st), &o), "convT2x2_dgrad");
  }
  return cuda_status(launch_fprop(BN, 0, a0, a0, b, p, num_tiles(p.g), Cin / BN, S(st)), "convT2x2_dgrad");
}

int clk_convT2x2_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout,
                       clk_stream_t st) {
  if (!x || !dy || !dw || N <= 0 || H <= 0 || W <= 0) return fail(CLK_E_BADARG, "convT2x2_wgrad: bad args");
  if (Cin % 64 || Cout % 64) return fail(CLK_E_UNSUPPORTED_SHAPE, "convT2x2_wgrad: channels must be multiples of 64");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  geom_quad(p.g, N, H, W, 64);
  taps_quad(p.g);
  p.ntaps = 4;
  p.G = 4;
  p.CU = Cin;
  p.CT = Cout;
  p.ct_split = Cout;
  p.m_tiles = (Cin + 127) / 128;
  p.n_tiles = Cout / 64;
  p.tap_groups = 1;
  p.tiles_total = num_tiles(p.g);
  p.ksplit = pick_ksplit(p.m_tiles * p.n_tiles, p.tiles_total);
  p.out = dw;
  p.ld_u = Cin;
  p.ld_t = Cout;
  CUtensorMap mu, mt;
  CHECK_RC(map_rows(&mu, x, N, H, W, Cin, p.g.tw, p.g.th));
  CHECK_RC(map_quad(&mt, dy, N, H, W, Cout, p.g.tw, p.g.th));
  return cuda_status(launch_wgrad(64, mu, mt, mt, p, S(st)), "convT2x2_wgrad");
}

// ------------------------------------------------------------------ BatchNorm / pool
#define REQ_C8(name, C) \
  if ((C) % 8 || (C) <= 0 || (C) > 2048) return fail(CLK_E_UNSUPPORTED_SHAPE, name ": C must be a multiple of 8 in (0, 2048] (C=%d)", (C))

int clk_bn_stats(const void* y, double* sum, double* sq, long long P, int C, clk_stream_t st) {
  if (!y || !sum || !sq || P <= 0) return fail(CLK_E_BADARG, "bn_stats: bad args");
  REQ_C8("bn_stats", C);
  return cuda_status(bn_stats(y, sum, sq, P, C, S(st)), "bn_stats");
}
int clk_bn_finalize(const double* sum, const double* sq, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float* mean_out, float* invstd_out, float* scale,
                    float* shift, int C, double count, float eps, float momentum, int training,
                    clk_stream_t st) {
  if (!gamma || !beta || !mean_out || !invstd_out || !scale || !shift || C <= 0)
    return fail(CLK_E_BADARG, "bn_finalize: bad args");
  if (training && (!sum || !sq || count <= 0)) return fail(CLK_E_BADARG, "bn_finalize: training needs sums");
  if (!training && (!running_mean || !running_var)) return fail(CLK_E_BADARG, "bn_finalize: eval needs running stats");
  return cuda_status(bn_finalize(sum, sq, gamma, beta, running_mean, running_var, mean_out, invstd_out, scale,
                                 shift, C, count, eps, momentum, training, S(st)),
                     "bn_finalize");
}
int clk_bn_apply(const void* y, void* z, const float* scale, const float* shift, long long P, int C,
                 clk_stream_t st) {
  if (!y || !z || !scale || !shift || P <= 0) return fail(CLK_E_BADARG, "bn_apply: bad args");
  REQ_C8("bn_apply", C);
  return cuda_status(bn_apply(y, z, scale, shift, P, C, S(st)), "bn_apply");
}
int clk_bn_apply_pool(const void* y, void* z, void* pooled, void* idx, const float* scale, const float* shift,
                      int N, int H, int W, int C, clk_stream_t st) {
  if (!y || !pooled || !idx || N <= 0) return fail(CLK_E_BADARG, "bn_apply_pool: bad args");
  if (scale && (!z || !shift)) return fail(CLK_E_BADARG, "bn_apply_pool: scale given without z/shift");
  if (H % 2 || W % 2) return fail(CLK_E_UNSUPPORTED_SHAPE, "bn_apply_pool: H and W must be even");
  REQ_C8("bn_apply_pool", C);
  return cuda_status(bn_apply_pool(y, z, pooled, idx, scale, shift, N, H, W, C, S(st)), "bn_apply_pool");
}
int clk_maxpool_bwd_add(const void* dpooled, const void* idx, const void* skip, void* din, int N, int H, int W,
                        int C, clk_stream_t st) {
  if (!dpooled || !idx || !din || N <= 0) return fail(CLK_E_BADARG, "maxpool_bwd_add: bad args");
  if (H % 2 || W % 2) return fail(CLK_E_UNSUPPORTED_SHAPE, "maxpool_bwd_add: H and W must be even");
  REQ_C8("maxpool_bwd_add", C);
  return cuda_status(maxpool_bwd_add(dpooled, idx, skip, din, N, H, W, C, S(st)), "maxpool_bwd_add");
}
int clk_maxpool_bwd_add_reduce(const void* dpooled, const void* idx, const void* skip, const void* y, void* din,
                               double* s1, double* s2, int N, int H, int W, int C, clk_stream_t st) {
  if (!dpooled || !idx || !din || !y || !s1 || !s2 || N <= 0) return fail(CLK_E_BADARG, "maxpool_bwd_add_reduce: bad args");
  if (H % 2 || W % 2) return fail(CLK_E_UNSUPPORTED_SHAPE, "maxpool_bwd_add_reduce: H and W must be even");
  REQ_C8("maxpool_bwd_add_reduce", C);
  if (C > 2048 || 256 % (C / 8) != 0)
    return fail(CLK_E_UNSUPPORTED_SHAPE, "maxpool_bwd_add_reduce: C / 8 must divide 256 (C=%d)", C);
  return cuda_status(maxpool_bwd_add_reduce(dpooled, idx, skip, y, din, s1, s2, N, H, W, C, S(st)),
                     "maxpool_bwd_add_reduce");
}
int clk_bn_bwd_reduce(const void* dz, const void* y, double* s1, double* s2, long long P, int C,
                      clk_stream_t st) {
  if (!dz || !y || !s1 || !s2 || P <= 0) return fail(CLK_E_BADARG, "bn_bwd_reduce: bad args");
  REQ_C8("bn_bwd_reduce", C);
  return cuda_status(bn_bwd_reduce(dz, y, s1, s2, P, C, S(st)), "bn_bwd_reduce");
}
int clk_bn_bwd_finalize(const double* s1, const double* s2, const float* gamma, const float* mean,
                        const float* invstd, float* dgamma, float* dbeta, float* kA, float* kB, float* kC, int C,
                        double count, int training, int accumulate, clk_stream_t st) {
  if (!s1 || !s2 || !gamma || !mean || !invstd || !dgamma || !dbeta || !kA || !kB || !kC || C <= 0 || count <= 0)
    return fail(CLK_E_BADARG, "bn_bwd_finalize: bad args");
  return cuda_status(bn_bwd_finalize(s1, s2, gamma, mean, invstd, dgamma, dbeta, kA, kB, kC, C, count, training,
                                     accumulate, S(st)),
                     "bn_bwd_finalize");
}
int clk_bn_relu_bwd_apply(const void* dz, const void* y, void* dpre, const float* kA, const float* kB,
                          const float* kC, double* dbias, long long P, int C, clk_stream_t st) {
  if (!dz || !y || !dpre || !kA || !kB || !kC || !dbias || P <= 0) return fail(CLK_E_BADARG, "bn_relu_bwd_apply: bad args");
  REQ_C8("bn_relu_bwd_apply", C);
  return cuda_status(bn_relu_bwd_apply(dz, y, dpre, kA, kB, kC, dbias, P, C, S(st)), "bn_relu_bwd_apply");
}
int clk_channel_sum(const void* g, double* out, long long P, int C, clk_stream_t st) {
  if (!g || !out || P <= 0) return fail(CLK_E_BADARG, "channel_sum: bad args");
  REQ_C8("channel_sum", C);
  return cuda_status(channel_sum(g, out, P, C, S(st)), "channel_sum");
}
int clk_f64_to_f32(const double* src, float* dst, int n, int ld_group, int groups, float alpha, int accumulate,
                   clk_stream_t st) {
  if (!src || !dst || n <= 0 || groups <= 0) return fail(CLK_E_BADARG, "f64_to_f32: bad args");
  return cuda_status(f64_to_f32(src, dst, n, ld_group, groups, alpha, accumulate, S(st)), "f64_to_f32");
}

// ------------------------------------------------------------------ loss / metrics / optimiser
int clk_ce_kd_loss(const float* logits, const float* old_logits, const int64_t* labels, long long P, int C,
                   int Cold, float T, float lambda, float gscale, void* dlogits, int ldd, double* loss_acc,
                   int* err_flag, clk_stream_t st) {
  if (!logits || !labels || !dlogits || !loss_acc || P <= 0 || C <= 0) return fail(CLK_E_BADARG, "ce_kd_loss: bad args");
  if (ldd % 8 || ldd < C) return fail(CLK_E_UNSUPPORTED_SHAPE, "ce_kd_loss: ldd must be a multiple of 8 and >= C");
  if (old_logits && (Cold <= 0 || Cold > C || T <= 0.f)) return fail(CLK_E_BADARG, "ce_kd_loss: bad distillation args");
  return cuda_status(ce_kd_loss(logits, old_logits, reinterpret_cast<const long long*>(labels), P, C, Cold, T,
                                lambda, gscale, dlogits, ldd, loss_acc, err_flag, S(st)),
                     "ce_kd_loss");
}
int clk_confusion_matrix(const int64_t* target, const int64_t* pred, long long n, int nc, int64_t* conf,
                         int* err_flag, clk_stream_t st) {
  if (!conf || nc <= 0 || n < 0) return fail(CLK_E_BADARG, "confusion_matrix: bad args");
  if (nc > 36) return fail(CLK_E_UNSUPPORTED_SHAPE, "confusion_matrix: nc must be <= 36 (nc=%d)", nc);
  if (n == 0) return CLK_OK;
  if (!target || !pred) return fail(CLK_E_BADARG, "confusion_matrix: null input");
  if ((reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(pred)) & 7)
    return fail(CLK_E_BADARG, "confusion_matrix: inputs must be 8-byte aligned int64 arrays");
  return cuda_status(confusion_matrix(reinterpret_cast<const long long*>(target),
                                      reinterpret_cast<const long long*>(pred), n, nc,
                                      reinterpret_cast<long long*>(conf), err_flag, S(st)),
                     "confusion_matrix");
}
int clk_argmax_confusion(const float* logits, const int64_t* labels, long long P, int C, int nc, int64_t* pred_out,
                         int64_t* conf, int64_t* correct, clk_stream_t st) {
  if (!logits || !labels || P <= 0 || C <= 0) return fail(CLK_E_BADARG, "argmax_confusion: bad args");
  if (conf && (nc < C || nc > 36)) return fail(CLK_E_BADARG, "argmax_confusion: need C <= nc <= 36 (a prediction >= nc has no bin)");
  if (C > 64) return fail(CLK_E_UNSUPPORTED_SHAPE, "argmax_confusion: C must be <= 64");
  return cuda_status(argmax_confusion(logits, reinterpret_cast<const long long*>(labels), P, C, conf ? nc : 1,
                                      reinterpret_cast<long long*>(pred_out), reinterpret_cast<long long*>(conf),
                                      reinterpret_cast<long long*>(correct), S(st)),
                     "argmax_confusion");
}
int clk_adam_multi_tensor(const void* tensors, const void* blocks, int nblocks, int chunk, float lr, float b1,
                          float b2, float eps, float bc1, float bc2_sqrt, float gscale, const float* hyper_dev,
                          clk_stream_t st) {
  if (!tensors || !blocks || nblocks < 0 || chunk <= 0 || chunk % 4) return fail(CLK_E_BADARG, "adam_multi_tensor: bad args");
  return cuda_status(adam_multi_tensor(static_cast<const AdamTensor*>(tensors), blocks, nblocks, chunk, lr, b1, b2,
                                       eps, bc1, bc2_sqrt, gscale, hyper_dev, S(st)),
                     "adam_multi_tensor");
}

#pragma GCC visibility pop
}  // extern "C"
