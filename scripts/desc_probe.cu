// desc_probe.cu — developer probe (not part of libclk): does a UMMA shared-memory descriptor whose start
// address is shifted by whole 128-byte rows inside a SWIZZLE_128B tile read the rows TMA wrote there?
// Tests base_offset = 0 vs base_offset = (start >> 7) & 7, for a K-major A operand and an MN-major B operand.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I continual_learning_b200/csrc -o /tmp/desc_probe scripts/desc_probe.cu && /tmp/desc_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "clk_ptx.cuh"

using namespace clk;

constexpr int ROWS = 288;  // rows of the big tile in smem (two TMA boxes of 144 rows)

__device__ __forceinline__ uint64_t desc_bo(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t bo) {
  return umma_smem_desc(saddr, lbo, sbo) | (static_cast<uint64_t>(bo & 7) << 49);
}

// mode 0: D[128 x 64] = A[shift .. shift+128][0:64] (K-major) * B[64 n][64 k]^T (K-major)
// mode 1: D[128 x 64] = U[64 k][128 m]^T (MN-major, 2 slabs) * T[shift .. shift+64][64 n] (MN-major, shifted in K)
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap mapBig,
                                             const __grid_constant__ CUtensorMap mapSmall, int mode, int shift,
                                             int sbo, int use_bo, float* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t s0 = smem_u32(raw);
  uint8_t* smem = raw + ((1024u - (s0 & 1023u)) & 1023u);
  uint8_t* sBig = smem;                       // ROWS x 128 B
  uint8_t* sSmall = smem + ROWS * 128;        // 128 x 128 B (two slabs of 64 rows in mode 1)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sSmall + 128 * 128);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, ROWS * 128 + 128 * 128);
    tma_load_3d(sBig, &mapBig, bar, 0, 0, 0);
    tma_load_3d(sBig + 144 * 128, &mapBig, bar, 0, 144, 0);
    if (mode == 0) {
      tma_load_3d(sSmall, &mapSmall, bar, 0, 0, 0);          // B: 64 rows (n) x 64 k  (+ 64 junk rows)
      tma_load_3d(sSmall + 64 * 128, &mapSmall, bar, 0, 0, 0);
    } else {
      tma_load_3d(sSmall, &mapSmall, bar, 0, 0, 0);           // U slab 0: 64 px x channels 0..63
      tma_load_3d(sSmall + 64 * 128, &mapSmall, bar, 64, 0, 0);  // U slab 1: channels 64..127
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t big = smem_u32(sBig), small = smem_u32(sSmall);
    if (mode == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      const uint32_t a0 = big + shift * 128;
      const uint32_t bo = use_bo ? ((a0 >> 7) & 7) : 0;
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem, desc_bo(a0 + k * 32, 16, sbo, bo), umma_smem_desc(small + k * 32, 16, 1024), idesc, k != 0);
    } else {
      const uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
      const uint32_t t0 = big + shift * 128;
      const uint32_t bo = use_bo ? ((t0 >> 7) & 7) : 0;
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem, umma_smem_desc(small + k * 2048, 64 * 128, 1024), desc_bo(t0 + k * 2 * sbo, 64 * 128, sbo, bo),
                  idesc, k != 0);
    }
    umma_commit(done);
  }
  __syncwarp();
  mbar_wait(done, 0);
  tc_fence_after();
  for (int chunk = 0; chunk < 2; ++chunk) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + chunk * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + chunk * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make3(EncodeTiledFn enc, void* base, uint64_t d0, uint64_t d1, uint32_t b0, uint32_t b1) {
  CUtensorMap m;
  cuuint64_t dims[3] = {d0, d1, 1};
  cuuint64_t st[2] = {d0 * 2, d0 * d1 * 2};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    exit(1);
  }
  return m;
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  std::vector<__nv_bfloat16> hBig(ROWS * 64), hB(64 * 64), hU(64 * 128);
  std::vector<float> fBig(ROWS * 64), fB(64 * 64), fU(64 * 128);
  srand(1);
  auto rnd = [] { return (float)((rand() % 17) - 8) / 8.0f; };
  for (int i = 0; i < ROWS * 64; ++i) { fBig[i] = rnd(); hBig[i] = __float2bfloat16(fBig[i]); }
  for (int i = 0; i < 64 * 64; ++i) { fB[i] = rnd(); hB[i] = __float2bfloat16(fB[i]); }
  for (int i = 0; i < 64 * 128; ++i) { fU[i] = rnd(); hU[i] = __float2bfloat16(fU[i]); }
  __nv_bfloat16 *dBig, *dB, *dU;
  float* dOut;
  cudaMalloc(&dBig, hBig.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dU, hU.size() * 2);
  cudaMalloc(&dOut, 128 * 64 * 4);
  cudaMemcpy(dBig, hBig.data(), hBig.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dU, hU.data(), hU.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap mBig = make3(enc, dBig, 64, ROWS, 64, 144);
  CUtensorMap mB = make3(enc, dB, 64, 64, 64, 64);
  CUtensorMap mU = make3(enc, dU, 128, 64, 64, 64);
  const int smem = ROWS * 128 + 128 * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> hOut(128 * 64);
  for (int mode = 0; mode < 2; ++mode)
    for (int sbo : {1024, 2048})
      for (int use_bo = 0; use_bo < 2; ++use_bo)
        for (int shift : {0, 1, 2, 3, 8, 9, 17}) {
          probe<<<1, 128, smem>>>(mBig, mode == 0 ? mB : mU, mode, shift, sbo, use_bo, dOut);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("mode %d shift %d: CUDA error %s\n", mode, shift, cudaGetErrorString(e));
            return 1;
          }
          cudaMemcpy(hOut.data(), dOut, hOut.size() * 4, cudaMemcpyDeviceToHost);
          double maxerr = 0;
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
              double ref = 0;
              if (mode == 0) {
                // row m of the window: 8-row groups are sbo bytes apart -> source row = shift + (m/8)*(sbo/128) + m%8
                const int src = shift + (m / 8) * (sbo / 128) + (m % 8);
                for (int k = 0; k < 64; ++k) ref += (double)fBig[src * 64 + k] * fB[n * 64 + k];
              } else {
                for (int k = 0; k < 64; ++k) {
                  const int src = shift + (k / 8) * (sbo / 128) + (k % 8);
                  ref += (double)fU[k * 128 + m] * fBig[src * 64 + n];
                }
              }
              maxerr = fmax(maxerr, fabs(ref - hOut[m * 64 + n]));
            }
          printf("mode %d (%s) sbo %4d base_offset %s shift %2d : max |err| = %.4f %s\n", mode,
                 mode == 0 ? "K-major A rows shifted" : "MN-major B k-rows shifted", sbo, use_bo ? "(addr>>7)&7" : "0         ",
                 shift, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
        }
  return 0;
}
