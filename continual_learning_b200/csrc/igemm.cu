// igemm.cu — tcgen05 / TMEM / TMA implicit-GEMM kernels (sm_100a only).
//
// Warp roles per CTA (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM
// allocator + single-thread tcgen05.mma issuer, warps 2..5 = epilogue (TMEM lane quadrant =
// warp % 4).  Operand tiles are [rows][64 bf16] with SWIZZLE_128B, written by TMA and read by the
// tensor core through UMMA shared-memory descriptors.  Accumulators live in TMEM (fp32).
//
// Replaces (reference): nn.Conv2d 3x3 / 1x1, nn.ConvTranspose2d 2x2 s2 forward and backward as
// dispatched by models/unet.py:13,16,28,31,34,50,53,66,69,72 and autograd (trainer.py:175).
#include "igemm.cuh"
#include <math_constants.h>

#include "clk_ptx.cuh"

namespace clk {

constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
static int g_fprop_sms = 148;
void igemm_set_num_sms(int n) { g_fprop_sms = n > 0 ? n : 148; }
// when a tensor-core kernel lets its successor be scheduled (PDL): 0 = at its very start, 1 = when a CTA's TMA producer
// has issued its last load (the successor's prologue overlaps the last tile's MMAs + epilogue), 2 = implicit at exit
__device__ int d_pdl_mode = 1;
cudaError_t igemm_set_pdl_mode(int mode) { return cudaMemcpyToSymbol(d_pdl_mode, &mode, sizeof(int)); }

// ------------------------------------------------------------------------------------------
// tile -> TMA base coordinates (coords 1..4; coord 0 is always the channel)
__device__ __forceinline__ void tile_base(const TileGeom& g, int tile, int& b1, int& b2, int& b3,
                                          int& b4) {
  if (g.mode == ADDR_LINEAR) {
    b1 = tile * g.tw;
    b2 = b3 = b4 = 0;
  } else if (g.mode == ADDR_NHWC) {
    const int tx = tile % g.tiles_w;
    const int r = tile / g.tiles_w;
    b1 = tx * g.tw;
    b2 = (r % g.tiles_h) * g.th;
    b3 = (r / g.tiles_h) * g.nb;
    b4 = 0;
  } else {  // ADDR_QUAD: coords {c, j, w, i, nh}
    const int tx = tile % g.tiles_w;
    b1 = 0;
    b2 = tx * g.tw;
    b3 = 0;
    b4 = (tile / g.tiles_w) * g.th;
  }
}

// row m of the tile -> linear pixel index of the tiled tensor (or -1 when outside)
// n, h, w are only produced when need_nhw (the ConvTranspose2d pixel-shuffle epilogue): the linear case would
// otherwise pay four integer divisions per thread and tile for nothing (P < 2^31 is checked by the callers)
__device__ __forceinline__ long long tile_row_pixel(const TileGeom& g, int b1, int b2, int b3,
                                                    int b4, int m, int& n, int& h, int& w, bool need_nhw = true) {
  if (g.mode == ADDR_LINEAR) {
    const unsigned int pix = static_cast<unsigned int>(b1) + static_cast<unsigned int>(m);
    if (static_cast<long long>(pix) >= static_cast<long long>(g.N) * g.H * g.W) return -1;
    if (need_nhw) {
      const unsigned int W = static_cast<unsigned int>(g.W), H = static_cast<unsigned int>(g.H);
      const unsigned int r = pix / W;
      w = static_cast<int>(pix - r * W);
      n = static_cast<int>(r / H);
      h = static_cast<int>(r - static_cast<unsigned int>(n) * H);
    }
    return pix;
  } else if (g.mode == ADDR_NHWC) {
    w = b1 + m % g.tw;
    const int r = m / g.tw;
    h = b2 + r % g.th;
    n = b3 + r / g.th;
    if (w >= g.W || h >= g.H || n >= g.N) return -1;
    return (static_cast<long long>(n) * g.H + h) * g.W + w;
  } else {
    w = b2 + m % g.tw;
    const int nh = b4 + m / g.tw;
    if (w >= g.W || nh >= g.N * g.H) return -1;
    n = nh / g.H;
    h = nh % g.H;
    return static_cast<long long>(nh) * g.W + w;
  }
}

__device__ __forceinline__ uint8_t* align1024(uint8_t* raw) {
  const uint32_t s = smem_u32(raw);
  return raw + ((1024u - (s & 1023u)) & 1023u);
}

// ------------------------------------------------------------------------------------------
// FPROP (generic): D[128 pixels, BN] = sum_{tap, chunk} A_tile(tap, chunk)[128 x 64] * B(tap, chunk)[BN x 64]^T
// Persistent: every CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; the TMA ring keeps running
// across tiles and two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
// Used for the 1-tap GEMMs (stem, 1x1 head, head dgrad, ConvTranspose2d forward), the 4-tap ConvTranspose2d
// dgrad and the conv3x3 layers on small feature maps (BN = 256).
constexpr int kFpScratch = 8 * 32 * 33 * 4;
constexpr int kFpThreads = 64 + 256;  // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quadrant)
constexpr int kFpEpiThreads = 256;

template <int BN, int STAGES, typename OutT>
__global__ void __launch_bounds__(kFpThreads, (BN <= 64 ? 2 : 1))
    igemm_fprop_kernel(const __grid_constant__ CUtensorMap mapA0,
                       const __grid_constant__ CUtensorMap mapA1,
                       const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapO,
                       const __grid_constant__ FpropParams p, const int m_tiles, const int n_tiles) {
  constexpr int A_BYTES = 128 * 128;
  constexpr int B_BYTES = BN * 128;
  constexpr uint32_t ACC_COLS = BN < 32 ? 32 : BN;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;

  if (d_pdl_mode == 0) pdl_launch_dependents();  // the next kernel of the stream may be scheduled and run its prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  float* scratch_all = reinterpret_cast<float*>(sB + STAGES * B_BYTES);
  float* s_bias = scratch_all + 8 * 32 * 33;
  float* s_sum = s_bias + BN;
  float* s_sq = s_sum + BN;
  uint64_t* full = reinterpret_cast<uint64_t*>(s_sq + BN);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kc = p.kc0 + p.kc1;
  const int num_kb = p.ntaps * kc;
  const int total = m_tiles * n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 8);
    }
    fence_mbar_init();
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 3 * BN; i += kFpThreads) s_bias[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // predecessor grid complete and flushed: inputs may be read, outputs written

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int n0 = (t / m_tiles) * BN;
        int b1, b2, b3, b4;
        tile_base(p.g, t % m_tiles, b1, b2, b3, b4);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], A_BYTES + B_BYTES);
          const int tap = kb / kc;
          const int ch = kb - tap * kc;
          const int c1 = b1 + p.g.t1[tap], c2 = b2 + p.g.t2[tap], c3 = b3 + p.g.t3[tap];
          if (ch < p.kc0)
            tma_load_5d(sA + s * A_BYTES, &mapA0, &full[s], ch * 64, c1, c2, c3, b4);
          else
            tma_load_5d(sA + s * A_BYTES, &mapA1, &full[s], (ch - p.kc0) * 64, c1, c2, c3, b4);
          tma_load_3d(sB + s * B_BYTES, &mapB, &full[s], ch * 64, n0, tap);
        }
      }
    }
    __syncwarp();
    if (d_pdl_mode == 1) pdl_launch_dependents();  // this CTA has requested its last operand tile
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (single thread)
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      uint32_t it = 0, itile = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++itile) {
        const uint32_t acc = itile & 1, pacc = (itile >> 1) & 1;
        mbar_wait(&acc_empty[acc], pacc ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * ACC_COLS;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a = smem_u32(sA + s * A_BYTES);
          const uint32_t b = smem_u32(sB + s * B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(d, umma_smem_desc(a + k * 32, 16, 1024), umma_smem_desc(b + k * 32, 16, 1024), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ epilogue: TMEM -> regs -> global
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;  // the two warps of a quadrant split the 32-column chunks
    const int m = q * 32 + lane;
    const int etid = threadIdx.x - 64;
    const bool do_stats = p.stat_sum != nullptr;
    const bool affine = p.bn_scale != nullptr;
    float* scratch = scratch_all + (warp - 2) * (32 * 33);
    uint8_t* stage_base = reinterpret_cast<uint8_t*>(scratch_all);  // two 16 KB staging tiles (TMA-store path)
    const bool tma_out = sizeof(OutT) == 2 && p.tma_store != 0;
    uint32_t nst = 0;  // staged 64-column groups so far (buffer parity)
    uint32_t itile = 0;
    int bias_nt = -1;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++itile) {
      const int nt = t / m_tiles;
      const int n0 = nt * BN;
      const uint32_t acc = itile & 1, pacc = (itile >> 1) & 1;
      if (nt != bias_nt) {
        named_bar_sync(2, kFpEpiThreads);
        for (int i = etid; i < BN; i += kFpEpiThreads) {
          if (do_stats && bias_nt >= 0) {  // flush the statistics of the N tile this CTA just left
            if (bias_nt * BN + i < p.n_store) {
              atomicAdd(&p.stat_sum[bias_nt * BN + i], static_cast<double>(s_sum[i]));
              atomicAdd(&p.stat_sq[bias_nt * BN + i], static_cast<double>(s_sq[i]));
            }
            s_sum[i] = 0.f;
            s_sq[i] = 0.f;
          }
          float bv = 0.f;
          if (p.bias != nullptr) {
            const int col = n0 + i;
            if (p.shuffle) bv = p.bias[col % p.cout_q];
            else if (col < p.n_store) bv = p.bias[col];
          }
          s_bias[i] = bv;
          if (affine) {  // inference: the statistics slots hold the BatchNorm scale / shift of this N tile
            s_sum[i] = n0 + i < p.n_store ? p.bn_scale[n0 + i] : 0.f;
            s_sq[i] = n0 + i < p.n_store ? p.bn_shift[n0 + i] : 0.f;
          }
        }
        named_bar_sync(2, kFpEpiThreads);
        bias_nt = nt;
      }
      int b1, b2, b3, b4;
      tile_base(p.g, t % m_tiles, b1, b2, b3, b4);
      int pn = 0, ph_ = 0, pw = 0;
      long long pix = tile_row_pixel(p.g, b1, b2, b3, b4, m, pn, ph_, pw, p.shuffle != 0);
      const bool valid = pix >= 0;
      OutT* dst = reinterpret_cast<OutT*>(p.dst0);
      int ld = p.ldc0;
      int colbase = n0;
      if (p.shuffle) {
        const int qd = n0 / p.cout_q;
        colbase = n0 - qd * p.cout_q;
        pix = (static_cast<long long>(pn) * (2 * p.g.H) + 2 * ph_ + (qd >> 1)) * (2 * p.g.W) + 2 * pw + (qd & 1);
      } else if (p.split_c > 0 && n0 >= p.split_c) {
        dst = reinterpret_cast<OutT*>(p.dst1);
        ld = p.ldc1;
        colbase = n0 - p.split_c;
      }
      OutT* drow = dst + (valid ? pix : 0) * ld + colbase;
      // ConvTranspose2d pixel shuffle with an N tile that spans several quadrants (BN > cout_q): the destination
      // pixel is chosen per 32-column chunk (cout_q is a multiple of 64, a chunk never straddles two quadrants)
      const bool shuffle_chunks = p.shuffle && BN > p.cout_q;
      const long long pix_q0 = (static_cast<long long>(pn) * (2 * p.g.H) + 2 * ph_) * (2 * p.g.W) + 2 * pw;

      mbar_wait(&acc_full[acc], pacc);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = half; chunk < BN / 32; chunk += 2) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * ACC_COLS + chunk * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v[j]) + s_bias[chunk * 32 + j];
          if (p.relu) x = fmaxf(x, 0.f);
          if (affine) x = fmaf(x, s_sum[chunk * 32 + j], s_sq[chunk * 32 + j]);
          f[j] = x;
        }
        if constexpr (sizeof(OutT) == 2) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
          if (tma_out) {
            // 64-column group i = chunk / 2: the two warps of a quadrant hold its two 32-column halves.  Stage the
            // [128 rows][128 B] tile in the SWIZZLE_128B pattern of the store map (conflict-free), then ONE TMA store
            // per group writes full lines: the per-thread stores below cost 32 LSU wave fronts per instruction
            // (one 16-byte piece per pixel row) and bound the short-K GEMMs (ConvTranspose2d, stem).
            uint8_t* sbuf = stage_base + (nst & 1) * 16384;
            if (etid == 0) tma_store_wait_read<1>();  // the store two groups back has read this buffer
            named_bar_sync(3, kFpEpiThreads);
            uint8_t* rp = sbuf + m * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(rp + ((((chunk & 1) * 4 + j) ^ (m & 7)) << 4)) =
                  make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            fence_proxy_async();
            named_bar_sync(3, kFpEpiThreads);
            const int gcol = n0 + (chunk >> 1) * 64;
            if (etid == 0 && gcol < p.n_store) {
              if (p.tma_store == 1) {
                tma_store_5d(&mapO, sbuf, gcol, b1, 0, 0, 0);
              } else if (p.tma_store == 2) {
                // the tile is 128 consecutive input pixels (raster order): rows b1 / W.., columns b1 % W.. of the
                // merged [N*H][W] grid; quadrant qd = (i, j) selects the output pixel (2h + i, 2w + j)
                const int qd = gcol / p.cout_q;
                tma_store_5d(&mapO, sbuf, gcol - qd * p.cout_q, qd & 1, b1 % p.g.W, qd >> 1, b1 / p.g.W);
              } else {
                tma_store_5d(&mapO, sbuf, gcol, b2, b4, 0, 0);
              }
              tma_store_commit();
            }
            ++nst;
          } else if (valid && (n0 + chunk * 32) < p.n_store) {
            OutT* crow = drow + chunk * 32;
            if (shuffle_chunks) {
              const int gcol = n0 + chunk * 32;
              const int qd = gcol / p.cout_q;
              crow = dst + (pix_q0 + (qd >> 1) * (2 * p.g.W) + (qd & 1)) * ld + (gcol - qd * p.cout_q);
            }
            uint4* o = reinterpret_cast<uint4*>(crow);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          }
          if (do_stats) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              f[2 * j] = valid ? bf16lo_to_f32(pk[j]) : 0.f;
              f[2 * j + 1] = valid ? bf16hi_to_f32(pk[j]) : 0.f;
            }
          }
        } else {
          const bool dense_rows = p.g.mode == ADDR_LINEAR && !p.shuffle && ld == p.n_store && BN == 32;
          if (dense_rows) {
            // fp32 rows of n_store (< 32) columns are contiguous in memory across the 32 consecutive pixels of a
            // warp: stage them in shared memory and write the whole 32*n_store block with coalesced stores
            const int ns = p.n_store;
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ns) scratch[lane * ns + j] = f[j];
            __syncwarp();
            OutT* wbase = dst + (static_cast<long long>(b1) + q * 32) * ld;  // first pixel row of this warp
            for (int idx = lane; idx < 32 * ns; idx += 32) {
              const int r = idx / ns;
              if ((vmask >> r) & 1u) wbase[idx] = scratch[idx];
            }
            __syncwarp();
          } else if (valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + chunk * 32 + j < p.n_store) drow[chunk * 32 + j] = f[j];
          }
          if (do_stats) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = valid ? f[j] : 0.f;
          }
        }
        if (do_stats) {
          float fq[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) fq[j] = f[j] * f[j];
          const float s = warp_colsum32(f, lane);
          const float sq = warp_colsum32(fq, lane);
          atomicAdd(&s_sum[chunk * 32 + lane], s);
          atomicAdd(&s_sq[chunk * 32 + lane], sq);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
    if (tma_out && etid == 0) tma_store_wait<0>();
    // statistics stay in shared memory across the tiles of one N tile: one fp64 atomic per channel per CTA and N tile
    if (do_stats && bias_nt >= 0) {
      named_bar_sync(1, kFpEpiThreads);
      for (int i = etid; i < BN; i += kFpEpiThreads) {
        if (bias_nt * BN + i < p.n_store) {
          atomicAdd(&p.stat_sum[bias_nt * BN + i], static_cast<double>(s_sum[i]));
          atomicAdd(&p.stat_sq[bias_nt * BN + i], static_cast<double>(s_sq[i]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------
// WGRAD: D[g][128 u, BN t] = sum_{pixel tiles} U_tile[64 px x 128 u]^T * T_tile(tap g)[64 px x BN t]
constexpr int kWgStagesMax = 8;
constexpr int kWgSlab = 64 * 128;  // [64 pixels][64 channels] bf16

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
    igemm_wgrad_kernel(const __grid_constant__ CUtensorMap mapU,
                       const __grid_constant__ CUtensorMap mapT0,
                       const __grid_constant__ CUtensorMap mapT1,
                       const __grid_constant__ WgradParams p, const int stages,
                       const uint32_t tmem_cols) {
  constexpr int NBS = BN / 64;  // T slabs per tap
  const int stage_bytes = (2 + p.G * NBS) * kWgSlab;

  if (d_pdl_mode == 0) pdl_launch_dependents();  // the next kernel of the stream may be scheduled and run its prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty = full + kWgStagesMax;
  uint64_t* tmem_full = empty + kWgStagesMax;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int bid = blockIdx.x;
  const int mt = bid % p.m_tiles;
  bid /= p.m_tiles;
  const int nt = bid % p.n_tiles;
  const int tg = bid / p.n_tiles;
  const int tap0 = tg * p.G;
  const int gcount = min(p.G, p.ntaps - tap0);
  const int cu0 = mt * 128;
  const int ct0 = nt * BN;
  const int per = (p.tiles_total + p.ksplit - 1) / p.ksplit;
  const int t_begin = blockIdx.y * per;
  const int num_kb = min(p.tiles_total, t_begin + per) - t_begin;
  if (num_kb <= 0) return;
  const int u_slabs = (cu0 + 64 < p.CU) ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&mapU);
    tma_prefetch_desc(&mapT0);
    tma_prefetch_desc(&mapT1);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // predecessor grid complete and flushed: inputs may be read, outputs written

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t tx_bytes = static_cast<uint32_t>(u_slabs + gcount * NBS) * kWgSlab;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], tx_bytes);
        int b1, b2, b3, b4;
        tile_base(p.g, t_begin + kb, b1, b2, b3, b4);
        uint8_t* sU = smem + s * stage_bytes;
        uint8_t* sT = sU + 2 * kWgSlab;
        // U is never shifted; in ADDR_QUAD it is the plain [N*H][W][C] view of the low-res tensor
        int u1 = b1, u2 = b2, u3 = b3, u4 = b4;
        if (p.g.mode == ADDR_QUAD) {
          u1 = b2;
          u2 = b4;
          u3 = 0;
          u4 = 0;
        }
        for (int i = 0; i < u_slabs; ++i)
          tma_load_5d(sU + i * kWgSlab, &mapU, &full[s], cu0 + i * 64, u1, u2, u3, u4);
        for (int g = 0; g < gcount; ++g) {
          const int tap = tap0 + g;
          const int c1 = b1 + p.g.t1[tap], c2 = b2 + p.g.t2[tap], c3 = b3 + p.g.t3[tap];
#pragma unroll
          for (int j = 0; j < NBS; ++j) {
            const int ct = ct0 + j * 64;
            uint8_t* d = sT + (g * NBS + j) * kWgSlab;
            if (ct < p.ct_split)
              tma_load_5d(d, &mapT0, &full[s], ct, c1, c2, c3, b4);
            else
              tma_load_5d(d, &mapT1, &full[s], ct - p.ct_split, c1, c2, c3, b4);
          }
        }
      }
    }
    __syncwarp();
    if (d_pdl_mode == 1) pdl_launch_dependents();  // this CTA has requested its last operand tile
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 1, 1);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % stages;
        const uint32_t ph = (kb / stages) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t u = smem_u32(smem + s * stage_bytes);
        const uint32_t t = u + 2 * kWgSlab;
        for (int g = 0; g < gcount; ++g) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(tmem_base + g * BN, umma_smem_desc(u + k * 2048, kWgSlab, 1024),
                      umma_smem_desc(t + g * NBS * kWgSlab + k * 2048, kWgSlab, 1024), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int u = cu0 + row;
    const bool valid = u < p.CU;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    for (int g = 0; g < gcount; ++g) {
      float* orow = p.out + (static_cast<size_t>(tap0 + g) * p.ld_u + (valid ? u : 0)) * p.ld_t + ct0;
      // transposed: out[tap][t][u] (the conv3x3 layout [tap][Cin][Cout]); lanes = consecutive u -> coalesced
      float* ocol = p.out + (static_cast<size_t>(tap0 + g) * p.ld_t + ct0) * p.ld_u + (valid ? u : 0);
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * BN + chunk * 32, v);
        tmem_ld_wait();
        if (valid) {
          if (p.transpose_out) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (ct0 + chunk * 32 + j < p.CT)
                atomicAdd(ocol + static_cast<size_t>(chunk * 32 + j) * p.ld_u, __uint_as_float(v[j]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (ct0 + chunk * 32 + 4 * j + 3 < p.CT)
                red_add_v4(orow + chunk * 32 + 4 * j, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                           __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------
// WGRAD9 (conv3x3, all nine taps per CTA, halo reuse).
//   D[pair p][row = 64*(tap - 2p) + ci][co] = sum_px X[px (+) tap][ci] * dY[px][co],  tap in {2p, 2p+1}
//   A operand = X halo tile (MN-major, M = 128 = two 64-channel "slabs" that are the SAME 64 channels
//   read through two different tap windows: LBO = byte distance between the two windows),
//   B operand = dY tile (MN-major, N = 64).  K = 128 pixels (16 rows x 8) per stage, 8 MMAs of K=16.
constexpr int kW9Stages = 4;
constexpr int kW9UBytes = 128 * 128;      // [16x8 px][64 ch]
constexpr int kW9TBytes = 18 * 16 * 128;  // [18 rows][16 px][64 ch]; row pitch 16 px = 2048 B
constexpr int kW9StageBytes = kW9UBytes + kW9TBytes;

__global__ void __launch_bounds__(kThreads, 1)
    igemm_wgrad9_kernel(const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapT0,
                        const __grid_constant__ CUtensorMap mapT1, const __grid_constant__ Wgrad9Params p) {
  if (d_pdl_mode == 0) pdl_launch_dependents();  // the next kernel of the stream may be scheduled and run its prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kW9Stages * kW9StageBytes);
  uint64_t* empty = full + kW9Stages;
  uint64_t* tmem_full = empty + kW9Stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int co_tile = blockIdx.x % p.cout_tiles;
  const int ci_slab = blockIdx.x / p.cout_tiles;
  const int per = (p.tiles_total + p.ksplit - 1) / p.ksplit;
  const int t_begin = blockIdx.y * per;
  const int num_kb = min(p.tiles_total, t_begin + per) - t_begin;
  if (num_kb <= 0) return;
  constexpr uint32_t TMEM_COLS = 512;  // 5 tap pairs x 64 columns

  if (threadIdx.x == 0) {
    for (int s = 0; s < kW9Stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&mapU);
    tma_prefetch_desc(&mapT0);
    tma_prefetch_desc(&mapT1);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // predecessor grid complete and flushed: inputs may be read, outputs written

  if (warp == 0) {
    if (elect_one()) {
      const bool src0 = ci_slab < p.split_slabs;
      const CUtensorMap* mapT = src0 ? &mapT0 : &mapT1;
      const int ct = (src0 ? ci_slab : ci_slab - p.split_slabs) * 64;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kW9Stages;
        const uint32_t ph = (kb / kW9Stages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], kW9StageBytes);
        const int tile = t_begin + kb;
        const int tx = tile % p.tiles_w;
        const int r = tile / p.tiles_w;
        const int ty = r % p.tiles_h;
        const int n = r / p.tiles_h;
        uint8_t* sU = smem + s * kW9StageBytes;
        tma_load_5d(sU, &mapU, &full[s], co_tile * 64, tx * 8, ty * 16, n, 0);
        tma_load_5d(sU + kW9UBytes, mapT, &full[s], ct, tx * 8 - 1, ty * 16 - 1, n, 0);
      }
    }
    __syncwarp();
    if (d_pdl_mode == 1) pdl_launch_dependents();  // this CTA has requested its last operand tile
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kW9Stages;
        const uint32_t ph = (kb / kW9Stages) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t u = smem_u32(smem + s * kW9StageBytes);
        const uint32_t t = u + kW9UBytes;
#pragma unroll
        for (int pr = 0; pr < 5; ++pr) {
          const int ta = 2 * pr, tb = (pr < 4) ? 2 * pr + 1 : 2 * pr;  // pair 4: tap 8 + a harmless shifted copy
          const uint32_t offa = ((ta / 3) * 16 + (ta % 3)) * 128;
          const uint32_t offb = (pr < 4) ? ((tb / 3) * 16 + (tb % 3)) * 128 : offa + 128;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            umma_bf16(tmem_base + pr * 64, umma_smem_desc(t + offa + k * 4096, offb - offa, 2048),
                      umma_smem_desc(u + k * 2048, 8192, 1024), idesc, (kb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ci = ci_slab * 64 + (row & 63);
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    for (int pr = 0; pr < 5; ++pr) {
      const int tap = 2 * pr + (row >> 6);
      // packed gradient layout [tap][Cin][Cout]: this thread owns one (tap, ci) row -> 32 consecutive floats per chunk
      float* obase = p.out + static_cast<size_t>(blockIdx.y) * p.split_stride +
                     (static_cast<size_t>(tap) * p.Cin + ci) * p.Cout + co_tile * 64;
#pragma unroll 1
      for (int chunk = 0; chunk < 2; ++chunk) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + pr * 64 + chunk * 32, v);
        tmem_ld_wait();
        if (tap < 9) {
          if (p.split_stride > 0) {  // this K split owns its own partial buffer: plain (deterministic) stores
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(obase + chunk * 32 + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              red_add_v4(obase + chunk * 32 + 4 * j, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

cudaError_t launch_wgrad9(const CUtensorMap& u, const CUtensorMap& t0, const CUtensorMap& t1,
                          const Wgrad9Params& p, cudaStream_t st) {
  const int smem = kW9Stages * kW9StageBytes + (2 * kW9Stages + 1) * 8 + 16 + 1024;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(igemm_wgrad9_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid(p.cin_slabs * p.cout_tiles, p.ksplit);
  launch_k(igemm_wgrad9_kernel, dim3(grid), dim3(kThreads), smem, st, u, t0, t1, p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// WGRAD9x2 (conv3x3 weight gradient, CTA pairs, cta_group::2, M = 256, N = 128).
//   A pair owns 64 input channels x 128 output channels x all nine taps.  Both CTAs load the SAME 64-channel X halo
//   tile, but CTA 1 loads it shifted down by one image row, so one shared A descriptor (start offset, LBO) selects
//   different taps in the two CTAs: three MMA groups x (2 CTAs x 2 row blocks) = 12 tap slots for the 9 taps:
//       group 0: offset (0,0), LBO = 1 px   -> CTA0 taps (0,0) (0,1)   CTA1 taps (1,0) (1,1)
//       group 1: offset (0,2), LBO = 2 rows -> CTA0 taps (0,2) (2,2)   CTA1 taps (1,2)  --
//       group 2: offset (2,0), LBO = 1 px   -> CTA0 taps (2,0) (2,1)   CTA1  --  --
//   The dY tile (B operand, MN-major, N = 128) is split 64 + 64 channels between the CTAs.  Per MMA and SM the
//   shared-memory port moves 6 KB per 64 cycles instead of 6 KB per 32 cycles in the 1-CTA kernel.
constexpr int kW2Stages = 4;
constexpr int kW2StageBytes = kW9UBytes + kW9TBytes;  // [128 px][64 cout] + [18 rows][16 px][64 cin]

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    igemm_wgrad9x2_kernel(const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapT0,
                          const __grid_constant__ CUtensorMap mapT1, const __grid_constant__ Wgrad9Params p) {
  if (d_pdl_mode == 0) pdl_launch_dependents();  // the next kernel of the stream may be scheduled and run its prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  // 2 KB pad after the last stage: CTA 1's unused tap slots read one halo row past their tile
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kW2Stages * kW2StageBytes + 2048);
  uint64_t* empty = full + kW2Stages;
  uint64_t* tmem_full = empty + kW2Stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cl = blockIdx.x >> 1;                 // cluster index within the (x) grid
  const int co_tile = cl % p.cout_tiles;          // 128-channel tiles of dY
  const int ci_slab = cl / p.cout_tiles;
  const int per = (p.tiles_total + p.ksplit - 1) / p.ksplit;
  const int t_begin = blockIdx.y * per;
  const int num_kb = min(p.tiles_total, t_begin + per) - t_begin;
  if (num_kb <= 0) return;  // uniform over the pair (same blockIdx.y)
  constexpr uint32_t TMEM_COLS = 512;  // 3 groups x 128 columns

  if (threadIdx.x == 0) {
    for (int s = 0; s < kW2Stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&mapU);
    tma_prefetch_desc(&mapT0);
    tma_prefetch_desc(&mapT1);
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // predecessor grid complete and flushed: inputs may be read, outputs written

  if (warp == 0) {
    if (elect_one()) {
      const bool src0 = ci_slab < p.split_slabs;
      const CUtensorMap* mapT = src0 ? &mapT0 : &mapT1;
      const int ct = (src0 ? ci_slab : ci_slab - p.split_slabs) * 64;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kW2Stages;
        const uint32_t ph = (kb / kW2Stages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        if (leader) mbar_arrive_expect_tx(&full[s], 2 * kW2StageBytes);
        const int tile = t_begin + kb;
        const int tx = tile % p.tiles_w;
        const int r = tile / p.tiles_w;
        const int ty = r % p.tiles_h;
        const int n = r / p.tiles_h;
        uint8_t* sU = smem + s * kW2StageBytes;
        const uint32_t lbar = leader_bar_addr(&full[s]);
        tma_load_5d_2sm(sU, &mapU, lbar, co_tile * 128 + 64 * static_cast<int>(rank), tx * 8, ty * 16, n, 0);
        tma_load_5d_2sm(sU + kW9UBytes, mapT, lbar, ct, tx * 8 - 1, ty * 16 - 1 + static_cast<int>(rank), n, 0);
      }
    }
    __syncwarp();
    if (d_pdl_mode == 1) pdl_launch_dependents();  // this CTA has requested its last operand tile
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, 128, 1, 1);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kW2Stages;
        const uint32_t ph = (kb / kW2Stages) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t u = smem_u32(smem + s * kW2StageBytes);
        const uint32_t t = u + kW9UBytes;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const uint32_t offa = g == 0 ? 0u : (g == 1 ? 2u * 128u : 2u * 2048u);
          const uint32_t lbo = g == 1 ? 2u * 2048u : 128u;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            umma_bf16_2cta(tmem_base + g * 128, umma_smem_desc(t + offa + k * 4096, lbo, 2048),
                           umma_smem_desc(u + k * 2048, 8192, 1024), idesc, (kb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit_2cta(&empty[s]);
      }
      umma_commit_2cta(tmem_full);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int slot = row >> 6;
    const int ci = ci_slab * 64 + (row & 63);
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int g = 0; g < 3; ++g) {
      // tap owned by (rank, group, slot); -1 = unused slot
      int tap;
      if (rank == 0) tap = g == 0 ? slot : (g == 1 ? (slot == 0 ? 2 : 8) : 6 + slot);
      else tap = g == 0 ? 3 + slot : (g == 1 && slot == 0 ? 5 : -1);
      float* obase = p.out + static_cast<size_t>(blockIdx.y) * p.split_stride +
                     (static_cast<size_t>(tap < 0 ? 0 : tap) * p.Cin + ci) * p.Cout + co_tile * 128;
#pragma unroll 1
      for (int chunk = 0; chunk < 4; ++chunk) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 128 + chunk * 32, v);
        tmem_ld_wait();
        if (tap >= 0) {
          if (p.split_stride > 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(obase + chunk * 32 + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              red_add_v4(obase + chunk * 32 + 4 * j, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
}

cudaError_t launch_wgrad9x2(const CUtensorMap& u, const CUtensorMap& t0, const CUtensorMap& t1,
                            const Wgrad9Params& p, cudaStream_t st) {
  const int smem = kW2Stages * kW2StageBytes + 2048 + (2 * kW2Stages + 1) * 8 + 16 + 1024;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(igemm_wgrad9x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid(2 * p.cin_slabs * p.cout_tiles, p.ksplit);  // cout_tiles = Cout / 128 here
  launch_k(igemm_wgrad9x2_kernel, dim3(grid), dim3(kThreads), smem, st, u, t0, t1, p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// CONV3 (conv3x3 forward / dgrad, halo variant, persistent).
//   super tile = 16 x 16 output pixels of one image = two M=128 sub tiles (columns 0-7 / 8-15);
//   A ring (2 stages): one [18 rows][24 px][64 ch] halo tile per 64-channel K chunk (row pitch 3072 B);
//   B ring (NB stages): one [BN][64] weight tile per (chunk, tap);
//   tap (r, s) of sub tile j reads the halo through a descriptor starting at ((r*24 + s + 8j) * 128) B,
//   SBO = 3072 (the hardware swizzle is a function of the absolute smem address, so whole-row shifts of
//   the start address stay consistent with what TMA wrote);
//   TMEM: 2 accumulator stages x 2 sub tiles x BN columns, epilogue of tile i overlaps main loop of i+1.
constexpr int kC3ABytes = 18 * 24 * 128;
constexpr int kC3Scratch = 8 * 32 * 33 * 4;
constexpr int kC3Threads = 64 + 256;  // TMA warp, MMA warp, 8 epilogue warps (4 per sub tile)
constexpr int kC3EpiThreads = 256;
template <int BN>
struct C3Cfg {
  static constexpr int NB = (BN == 128) ? 5 : 8;
  static constexpr int BBytes = BN * 128;
  static constexpr int Smem = 2 * kC3ABytes + NB * BBytes + kC3Scratch + 3 * BN * 4 + (8 + 2 * NB) * 8 + 16 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(kC3Threads, 1)
    igemm_conv3_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                       const __grid_constant__ CUtensorMap mapB, const __grid_constant__ Conv3Params p) {
  constexpr int NB = C3Cfg<BN>::NB;
  constexpr int B_BYTES = C3Cfg<BN>::BBytes;
  constexpr uint32_t TMEM_COLS = 4 * BN;

  if (d_pdl_mode == 0) pdl_launch_dependents();  // the next kernel of the stream may be scheduled and run its prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * kC3ABytes;
  float* scratch_all = reinterpret_cast<float*>(sB + NB * B_BYTES);
  float* s_sum = scratch_all + 8 * 32 * 33;
  float* s_sq = s_sum + BN;
  float* s_bias = s_sq + BN;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(s_bias + BN);
  uint64_t* a_empty = a_full + 2;
  uint64_t* acc_full = a_empty + 2;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* b_full = acc_empty + 2;
  uint64_t* b_empty = b_full + NB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + NB);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kc = p.kc0 + p.kc1;
  const int total = p.m_tiles * p.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 8);
    }
    for (int s = 0; s < NB; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 3 * BN; i += kC3Threads) s_sum[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // predecessor grid complete and flushed: inputs may be read, outputs written

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    // Two cursors over the flattened (tile, chunk) items: the halo cursor (A ring, 2 stages) and the weight
    // cursor (B ring, 9 tiles per item).  Both are advanced by non-blocking polls, so the halo tile of item i+1 is
    // requested the moment item i-1 releases its stage — a full item (9*8 MMAs) ahead of its first use — instead
    // of queueing behind the nine weight loads of item i.
    if (elect_one()) {
      uint32_t ia = 0, ib = 0;                 // A loads / B loads issued so far
      int tA = blockIdx.x, chA = 0;            // next halo tile to request
      int tB = blockIdx.x, chB = 0, tapB = 0;  // next weight tile to request
      uint32_t idle = 0;
      while (tB < total) {
        if (++idle > CLK_SPIN_LIMIT) __trap();  // protocol bug: fail the launch instead of hanging the GPU
        if (tA < total && ia <= ib / 9 + 1) {
          const uint32_t sa = ia & 1, pa = (ia >> 1) & 1;
          if (mbar_try_wait(&a_empty[sa], pa ^ 1)) {
            const int m = tA % p.m_tiles;
            const int tx = m % p.tiles_w;
            const int r = m / p.tiles_w;
            const int ty = r % p.tiles_h;
            const int n = r / p.tiles_h;
            mbar_arrive_expect_tx(&a_full[sa], kC3ABytes);
            if (chA < p.kc0)
              tma_load_5d(sA + sa * kC3ABytes, &mapA0, &a_full[sa], chA * 64, tx * 16 - 1, ty * 16 - 1, n, 0);
            else
              tma_load_5d(sA + sa * kC3ABytes, &mapA1, &a_full[sa], (chA - p.kc0) * 64, tx * 16 - 1, ty * 16 - 1, n, 0);
            ++ia;
            idle = 0;
            if (++chA == kc) {
              chA = 0;
              tA += gridDim.x;
            }
          }
        }
        {
          const uint32_t sb = ib % NB, pb = (ib / NB) & 1;
          if (mbar_try_wait(&b_empty[sb], pb ^ 1)) {
            mbar_arrive_expect_tx(&b_full[sb], B_BYTES);
            tma_load_3d(sB + sb * B_BYTES, &mapB, &b_full[sb], chB * 64, (tB / p.m_tiles) * BN, tapB);
            ++ib;
            idle = 0;
            if (++tapB == 9) {
              tapB = 0;
              if (++chB == kc) {
                chB = 0;
                tB += gridDim.x;
              }
            }
          }
        }
      }
    }
    __syncwarp();
    if (d_pdl_mode == 1) pdl_launch_dependents();  // this CTA has requested its last operand tile
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      uint32_t ia = 0, ib = 0, it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const uint32_t acc = it & 1, pacc = (it >> 1) & 1;
        mbar_wait(&acc_empty[acc], pacc ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * 2 * BN;
        for (int ch = 0; ch < kc; ++ch) {
          const uint32_t sa = ia & 1, pa = (ia >> 1) & 1;
          mbar_wait(&a_full[sa], pa);
          const uint32_t abase = smem_u32(sA + sa * kC3ABytes);
#pragma unroll 1
          for (int tr = 0; tr < 3; ++tr) {
#pragma unroll 1
            for (int tsx = 0; tsx < 3; ++tsx) {
              const uint32_t sb = ib % NB, pb = (ib / NB) & 1;
              mbar_wait(&b_full[sb], pb);
              tc_fence_after();
              const uint32_t a = abase + (tr * 24 + tsx) * 128;
              const uint32_t b = smem_u32(sB + sb * B_BYTES);
#pragma unroll
              for (int j = 0; j < 2; ++j) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_bf16(d0 + j * BN, umma_smem_desc(a + j * 1024 + k * 32, 16, 3072),
                            umma_smem_desc(b + k * 32, 16, 1024), idesc, (ch | tr | tsx | k) != 0 ? 1u : 0u);
                }
              }
              umma_commit(&b_empty[sb]);
              ++ib;
            }
          }
          umma_commit(&a_empty[sa]);
          ++ia;
        }
        umma_commit(&acc_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ epilogue (warps 2..9: quadrant = warp % 4, sub tile = (warp-2)/4)
    const int q = warp & 3;
    const int j = (warp - 2) >> 2;
    const int mrow = q * 32 + lane;
    const int etid = threadIdx.x - 64;
    const bool do_stats = p.stat_sum != nullptr;
    const bool affine = p.bn_scale != nullptr;
    float* scratch = scratch_all + (warp - 2) * (32 * 33);
    uint32_t it = 0;
    int bias_nt = -1;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int m = t % p.m_tiles, nt = t / p.m_tiles;
      const int tx = m % p.tiles_w;
      const int r = m / p.tiles_w;
      const int ty = r % p.tiles_h;
      const int n = r / p.tiles_h;
      const int n0 = nt * BN;
      const uint32_t acc = it & 1, pacc = (it >> 1) & 1;
      if (nt != bias_nt) {  // (uniform across the epilogue warps) stage this N tile's bias in smem
        named_bar_sync(2, kC3EpiThreads);
        for (int i = etid; i < BN; i += kC3EpiThreads) {
          if (do_stats && bias_nt >= 0) {  // flush the statistics of the N tile this CTA just left
            if (bias_nt * BN + i < p.n_store) {
              atomicAdd(&p.stat_sum[bias_nt * BN + i], static_cast<double>(s_sum[i]));
              atomicAdd(&p.stat_sq[bias_nt * BN + i], static_cast<double>(s_sq[i]));
            }
            s_sum[i] = 0.f;
            s_sq[i] = 0.f;
          }
          s_bias[i] = (p.bias != nullptr && n0 + i < p.n_store) ? p.bias[n0 + i] : 0.f;
          if (affine) {  // inference: the statistics slots hold the BatchNorm scale / shift of this N tile
            s_sum[i] = n0 + i < p.n_store ? p.bn_scale[n0 + i] : 0.f;
            s_sq[i] = n0 + i < p.n_store ? p.bn_shift[n0 + i] : 0.f;
          }
        }
        named_bar_sync(2, kC3EpiThreads);
        bias_nt = nt;
      }
      const int h = ty * 16 + (mrow >> 3);
      const int w = tx * 16 + j * 8 + (mrow & 7);
      const bool valid = h < p.H && w < p.W;
      const long long pixoff = valid ? (static_cast<long long>(n) * p.H + h) * p.W + w : 0;
      mbar_wait(&acc_full[acc], pacc);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 2 * BN + j * BN + chunk * 32, v);
        tmem_ld_wait();
        const bool in_store = (n0 + chunk * 32) < p.n_store;
        uint32_t pk[16];
        const float4* b4 = reinterpret_cast<const float4*>(s_bias + chunk * 32);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const float4 bb = b4[jj];
          float x0 = __uint_as_float(v[4 * jj]) + bb.x, x1 = __uint_as_float(v[4 * jj + 1]) + bb.y;
          float x2 = __uint_as_float(v[4 * jj + 2]) + bb.z, x3 = __uint_as_float(v[4 * jj + 3]) + bb.w;
          if (p.relu) {
            x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f);
          }
          if (affine) {
            const float4 sc = reinterpret_cast<const float4*>(s_sum + chunk * 32)[jj];
            const float4 sh = reinterpret_cast<const float4*>(s_sq + chunk * 32)[jj];
            x0 = fmaf(x0, sc.x, sh.x); x1 = fmaf(x1, sc.y, sh.y); x2 = fmaf(x2, sc.z, sh.z); x3 = fmaf(x3, sc.w, sh.w);
          }
          pk[2 * jj] = pack_bf16x2(x0, x1);
          pk[2 * jj + 1] = pack_bf16x2(x2, x3);
        }
        if (valid && in_store) {
          // two destinations (dgrad of a concat input): the split is a multiple of 32 columns, so it is decided
          // per 32-column chunk and an N tile may straddle it
          const int gc = n0 + chunk * 32;
          __nv_bfloat16* orow = (p.split_c > 0 && gc >= p.split_c)
                                    ? reinterpret_cast<__nv_bfloat16*>(p.dst1) + pixoff * p.ldc1 + (gc - p.split_c)
                                    : reinterpret_cast<__nv_bfloat16*>(p.dst0) + pixoff * p.ldc0 + gc;
          uint4* o = reinterpret_cast<uint4*>(orow);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            o[jj] = make_uint4(pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
        }
        if (do_stats) {
          // statistics of the values exactly as stored (bf16-rounded), reduced over the 32 rows with shuffles
          float fs[32], fq[32];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const float lo = valid ? bf16lo_to_f32(pk[jj]) : 0.f, hi = valid ? bf16hi_to_f32(pk[jj]) : 0.f;
            fs[2 * jj] = lo;
            fs[2 * jj + 1] = hi;
            fq[2 * jj] = lo * lo;
            fq[2 * jj + 1] = hi * hi;
          }
          const float s = warp_colsum32(fs, lane);
          const float sq = warp_colsum32(fq, lane);
          atomicAdd(&s_sum[chunk * 32 + lane], s);
          atomicAdd(&s_sq[chunk * 32 + lane], sq);
        }
      }
      // accumulator stage drained: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
    // statistics stay in shared memory across the tiles of one N tile: one fp64 atomic per channel per CTA and N tile
    if (do_stats && bias_nt >= 0) {
      named_bar_sync(1, kC3EpiThreads);
      for (int i = etid; i < BN; i += kC3EpiThreads) {
        if (bias_nt * BN + i < p.n_store) {
          atomicAdd(&p.stat_sum[bias_nt * BN + i], static_cast<double>(s_sum[i]));
          atomicAdd(&p.stat_sq[bias_nt * BN + i], static_cast<double>(s_sq[i]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN>
static cudaError_t launch_conv3_t(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                                  const Conv3Params& p, int num_sms, cudaStream_t st) {
  constexpr int smem = C3Cfg<BN>::Smem;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(igemm_conv3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  int grid = p.m_tiles * p.n_tiles;
  if (grid > num_sms) grid = num_sms;
  launch_k(igemm_conv3_kernel<BN>, dim3(grid), dim3(kC3Threads), smem, st, a0, a1, b, p);
  return cudaGetLastError();
}

cudaError_t launch_conv3(int BN, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                         const Conv3Params& p, int num_sms, cudaStream_t st) {
  if (BN == 128) return launch_conv3_t<128>(a0, a1, b, p, num_sms, st);
  if (BN == 64) return launch_conv3_t<64>(a0, a1, b, p, num_sms, st);
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------
// CONV3x2 (conv3x3 forward / dgrad, halo variant, CTA PAIRS: tcgen05.mma.cta_group::2, M = 256).
//   A cluster of two CTAs owns one 16x16-pixel super tile; CTA r owns the 16x8 half r (its own halo tile,
//   18 rows x 16 px, row pitch 2048 B) = 128 rows of the M=256 MMA and 128 lanes of accumulator in its own TMEM;
//   the [BN][64] weight tile of a tap is split: each CTA loads BN/2 rows, the tensor cores of the pair share them.
//   Per MMA and per SM the shared-memory port then moves A 4 KB + B BN*16 B instead of A 4 KB + B BN*32 B, and each
//   weight byte is written to shared memory once per PAIR — the port, not the tensor pipe, limits the 1-CTA kernels.
//   Leader (rank 0): issues all MMAs, owns the full barriers (both CTAs' TMA loads credit them) and the
//   accumulator-empty barriers (16 arrivals: 8 epilogue warps of each CTA); empty / accumulator-full barriers live in
//   both CTAs and are signalled by multicast tcgen05.commit.
// SUB = number of 16x8 sub tiles per CTA: 1 for BN = 256 (the pair covers 16x16 pixels), 2 for BN <= 128 (the pair
// covers 16x32 pixels, every weight tile then feeds four M=128 row blocks).
template <int BN, int SUB>
struct X2Cfg {
  static constexpr int PitchPx = SUB == 1 ? 16 : 24;           // halo row pitch in pixels (box width)
  static constexpr int ABytes = 18 * PitchPx * 128;
  static constexpr int AS = SUB == 1 ? 3 : 2;                  // halo stages
  static constexpr int NB = (BN == 256) ? 6 : 8;               // weight stages
  static constexpr int BBytes = (BN / 2) * 128;
  static constexpr int Smem = AS * ABytes + NB * BBytes + 3 * BN * 4 + (2 * AS + 2 * NB + 4) * 8 + 16 + 1024;
};

template <int BN, int SUB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kC3Threads, 1)
    igemm_conv3x2_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                         const __grid_constant__ CUtensorMap mapB, const __grid_constant__ Conv3Params p) {
  using Cfg = X2Cfg<BN, SUB>;
  constexpr int NB = Cfg::NB;
  constexpr int B_BYTES = Cfg::BBytes;
  constexpr uint32_t TMEM_COLS = 2 * SUB * BN;
  constexpr int AS = Cfg::AS;
  constexpr int kX2ABytes = Cfg::ABytes;
  constexpr int PITCH = Cfg::PitchPx;       // pixels
  constexpr int TILE_W = 8 * SUB;           // output columns owned by one CTA

  if (d_pdl_mode == 0) pdl_launch_dependents();  // the next kernel of the stream may be scheduled and run its prologue
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + AS * kX2ABytes;
  float* s_sum = reinterpret_cast<float*>(sB + NB * B_BYTES);
  float* s_sq = s_sum + BN;
  float* s_bias = s_sq + BN;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(s_bias + BN);
  uint64_t* a_empty = a_full + AS;
  uint64_t* acc_full = a_empty + AS;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* b_full = acc_empty + 2;
  uint64_t* b_empty = b_full + NB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_empty + NB);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int nclusters = gridDim.x >> 1;
  const int kc = p.kc0 + p.kc1;
  const int total = p.m_tiles * p.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < AS; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 16);
    }
    for (int s = 0; s < NB; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  for (int i = threadIdx.x; i < 3 * BN; i += kC3Threads) s_sum[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / cross-CTA TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // predecessor grid complete and flushed: inputs may be read, outputs written

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (one per CTA)
    if (elect_one()) {
      uint32_t ia = 0, ib = 0;
      int tA = cluster_id, chA = 0;
      int tB = cluster_id, chB = 0, tapB = 0;
      uint32_t idle = 0;
      while (tB < total) {
        if (++idle > CLK_SPIN_LIMIT) __trap();
        if (tA < total && ia <= ib / 9 + (AS - 1)) {
          const uint32_t sa = ia % AS, pa = (ia / AS) & 1;
          if (mbar_try_wait(&a_empty[sa], pa ^ 1)) {
            const int m = tA % p.m_tiles;
            const int tx = m % p.tiles_w;
            const int r = m / p.tiles_w;
            const int ty = r % p.tiles_h;
            const int n = r / p.tiles_h;
            if (leader) mbar_arrive_expect_tx(&a_full[sa], 2 * kX2ABytes);
            const int w0 = tx * (2 * TILE_W) + TILE_W * static_cast<int>(rank) - 1;
            if (chA < p.kc0)
              tma_load_5d_2sm(sA + sa * kX2ABytes, &mapA0, leader_bar_addr(&a_full[sa]), chA * 64, w0, ty * 16 - 1, n, 0);
            else
              tma_load_5d_2sm(sA + sa * kX2ABytes, &mapA1, leader_bar_addr(&a_full[sa]), (chA - p.kc0) * 64, w0,
                              ty * 16 - 1, n, 0);
            ++ia;
            idle = 0;
            if (++chA == kc) {
              chA = 0;
              tA += nclusters;
            }
          }
        }
        {
          const uint32_t sb = ib % NB, pb = (ib / NB) & 1;
          if (mbar_try_wait(&b_empty[sb], pb ^ 1)) {
            if (leader) mbar_arrive_expect_tx(&b_full[sb], 2 * B_BYTES);
            tma_load_3d_2sm(sB + sb * B_BYTES, &mapB, leader_bar_addr(&b_full[sb]), chB * 64,
                            (tB / p.m_tiles) * BN + static_cast<int>(rank) * (BN / 2), tapB);
            ++ib;
            idle = 0;
            if (++tapB == 9) {
              tapB = 0;
              if (++chB == kc) {
                chB = 0;
                tB += nclusters;
              }
            }
          }
        }
      }
    }
    __syncwarp();
    if (d_pdl_mode == 1) pdl_launch_dependents();  // this CTA has requested its last operand tile
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer: leader CTA only
    if (leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
      uint32_t ia = 0, ib = 0, it = 0;
      for (int t = cluster_id; t < total; t += nclusters, ++it) {
        const uint32_t acc = it & 1, pacc = (it >> 1) & 1;
        mbar_wait(&acc_empty[acc], pacc ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * SUB * BN;
        for (int ch = 0; ch < kc; ++ch) {
          const uint32_t sa = ia % AS, pa = (ia / AS) & 1;
          mbar_wait(&a_full[sa], pa);
          // all operand tiles lie below 256 KB, so moving a descriptor inside its tile is ONE add on the 14-bit
          // start-address field (>> 4): the single issuing thread stays well ahead of the tensor pipe
          const uint64_t adesc0 = umma_smem_desc(smem_u32(sA + sa * kX2ABytes), 16, PITCH * 128);
#pragma unroll 1
          for (int tr = 0; tr < 3; ++tr) {
#pragma unroll 1
            for (int tsx = 0; tsx < 3; ++tsx) {
              const uint32_t sb = ib % NB, pb = (ib / NB) & 1;
              mbar_wait(&b_full[sb], pb);
              tc_fence_after();
              const uint64_t adesc = adesc0 + static_cast<uint64_t>(((tr * PITCH + tsx) * 128) >> 4);
              const uint64_t bdesc = umma_smem_desc(smem_u32(sB + sb * B_BYTES), 16, 1024);
#pragma unroll
              for (int j = 0; j < SUB; ++j) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_bf16_2cta(d0 + j * BN, adesc + static_cast<uint64_t>((j * 1024 + k * 32) >> 4),
                                 bdesc + static_cast<uint64_t>((k * 32) >> 4), idesc, (ch | tr | tsx | k) != 0 ? 1u : 0u);
                }
              }
              umma_commit_2cta(&b_empty[sb]);
              ++ib;
            }
          }
          umma_commit_2cta(&a_empty[sa]);
          ++ia;
        }
        umma_commit_2cta(&acc_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ epilogue (8 warps per CTA: quadrant = warp % 4, the two warps
    // of a quadrant split the 32-column chunks)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    // SUB == 1: the two warps of a quadrant split the column chunks; SUB == 2: they take one sub tile each
    const int j = SUB == 2 ? half : 0;
    const int chunk0 = SUB == 2 ? 0 : half;
    const int chunk_step = SUB == 2 ? 1 : 2;
    const int mrow = q * 32 + lane;
    const int etid = threadIdx.x - 64;
    const bool do_stats = p.stat_sum != nullptr;
    const bool affine = p.bn_scale != nullptr;
    uint32_t it = 0;
    int bias_nt = -1;
    for (int t = cluster_id; t < total; t += nclusters, ++it) {
      const int m = t % p.m_tiles, nt = t / p.m_tiles;
      const int tx = m % p.tiles_w;
      const int r = m / p.tiles_w;
      const int ty = r % p.tiles_h;
      const int n = r / p.tiles_h;
      const int n0 = nt * BN;
      const uint32_t acc = it & 1, pacc = (it >> 1) & 1;
      if (nt != bias_nt) {
        named_bar_sync(2, kC3EpiThreads);
        for (int i = etid; i < BN; i += kC3EpiThreads) {
          if (do_stats && bias_nt >= 0) {  // flush the statistics of the N tile this CTA just left
            if (bias_nt * BN + i < p.n_store) {
              atomicAdd(&p.stat_sum[bias_nt * BN + i], static_cast<double>(s_sum[i]));
              atomicAdd(&p.stat_sq[bias_nt * BN + i], static_cast<double>(s_sq[i]));
            }
            s_sum[i] = 0.f;
            s_sq[i] = 0.f;
          }
          s_bias[i] = (p.bias != nullptr && n0 + i < p.n_store) ? p.bias[n0 + i] : 0.f;
          if (affine) {  // inference: the statistics slots hold the BatchNorm scale / shift of this N tile
            s_sum[i] = n0 + i < p.n_store ? p.bn_scale[n0 + i] : 0.f;
            s_sq[i] = n0 + i < p.n_store ? p.bn_shift[n0 + i] : 0.f;
          }
        }
        named_bar_sync(2, kC3EpiThreads);
        bias_nt = nt;
      }
      const int h = ty * 16 + (mrow >> 3);
      const int w = tx * (2 * TILE_W) + TILE_W * static_cast<int>(rank) + j * 8 + (mrow & 7);
      const bool valid = h < p.H && w < p.W;
      const long long pixoff = valid ? (static_cast<long long>(n) * p.H + h) * p.W + w : 0;
      mbar_wait(&acc_full[acc], pacc);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = chunk0; chunk < BN / 32; chunk += chunk_step) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * SUB * BN + j * BN + chunk * 32, v);
        tmem_ld_wait();
        const bool in_store = (n0 + chunk * 32) < p.n_store;
        uint32_t pk[16];
        const float4* b4 = reinterpret_cast<const float4*>(s_bias + chunk * 32);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const float4 bb = b4[jj];
          float x0 = __uint_as_float(v[4 * jj]) + bb.x, x1 = __uint_as_float(v[4 * jj + 1]) + bb.y;
          float x2 = __uint_as_float(v[4 * jj + 2]) + bb.z, x3 = __uint_as_float(v[4 * jj + 3]) + bb.w;
          if (p.relu) {
            x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f);
          }
          if (affine) {
            const float4 sc = reinterpret_cast<const float4*>(s_sum + chunk * 32)[jj];
            const float4 sh = reinterpret_cast<const float4*>(s_sq + chunk * 32)[jj];
            x0 = fmaf(x0, sc.x, sh.x); x1 = fmaf(x1, sc.y, sh.y); x2 = fmaf(x2, sc.z, sh.z); x3 = fmaf(x3, sc.w, sh.w);
          }
          pk[2 * jj] = pack_bf16x2(x0, x1);
          pk[2 * jj + 1] = pack_bf16x2(x2, x3);
        }
        if (valid && in_store) {
          const int gc = n0 + chunk * 32;
          __nv_bfloat16* orow = (p.split_c > 0 && gc >= p.split_c)
                                    ? reinterpret_cast<__nv_bfloat16*>(p.dst1) + pixoff * p.ldc1 + (gc - p.split_c)
                                    : reinterpret_cast<__nv_bfloat16*>(p.dst0) + pixoff * p.ldc0 + gc;
          uint4* o = reinterpret_cast<uint4*>(orow);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            o[jj] = make_uint4(pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
        }
        if (do_stats) {
          float fs[32], fq[32];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const float lo = valid ? bf16lo_to_f32(pk[jj]) : 0.f, hi = valid ? bf16hi_to_f32(pk[jj]) : 0.f;
            fs[2 * jj] = lo;
            fs[2 * jj + 1] = hi;
            fq[2 * jj] = lo * lo;
            fq[2 * jj + 1] = hi * hi;
          }
          const float s = warp_colsum32(fs, lane);
          const float sq = warp_colsum32(fq, lane);
          atomicAdd(&s_sum[chunk * 32 + lane], s);
          atomicAdd(&s_sq[chunk * 32 + lane], sq);
        }
      }
      // this warp is done with the accumulator stage: tell the leader's MMA thread (16 arrivals free the stage)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&acc_empty[acc], 0);
    }
    // statistics stay in shared memory across the tiles of one N tile: one fp64 atomic per channel per CTA and N tile
    if (do_stats && bias_nt >= 0) {
      named_bar_sync(1, kC3EpiThreads);
      for (int i = etid; i < BN; i += kC3EpiThreads) {
        if (bias_nt * BN + i < p.n_store) {
          atomicAdd(&p.stat_sum[bias_nt * BN + i], static_cast<double>(s_sum[i]));
          atomicAdd(&p.stat_sq[bias_nt * BN + i], static_cast<double>(s_sq[i]));
        }
      }
    }
  }

  // the leader's MMAs read the peer's shared memory and write its TMEM: nobody leaves before everybody is done
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
}

template <int BN, int SUB>
static cudaError_t launch_conv3x2_t(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                                    const Conv3Params& p, int num_sms, cudaStream_t st) {
  constexpr int smem = X2Cfg<BN, SUB>::Smem;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(igemm_conv3x2_kernel<BN, SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  int clusters = p.m_tiles * p.n_tiles;
  if (clusters > num_sms / 2) clusters = num_sms / 2;
  launch_k(igemm_conv3x2_kernel<BN, SUB>, dim3(2 * clusters), dim3(kC3Threads), smem, st, a0, a1, b, p);
  return cudaGetLastError();
}

// BN = 256: one sub tile per CTA (a0/a1 box {64, 16, 18}, tiles of 16x16 px per pair);
// BN = 128 / 64: two sub tiles per CTA (box {64, 24, 18}, tiles of 16x32 px per pair).  b box {64, BN/2, 1}.
// BN = 128 with one sub tile: for layers with so few 256-column tiles that half of the clusters would idle.
cudaError_t launch_conv3x2(int BN, int SUB, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                           const Conv3Params& p, int num_sms, cudaStream_t st) {
  if (BN == 256 && SUB == 1) return launch_conv3x2_t<256, 1>(a0, a1, b, p, num_sms, st);
  if (BN == 128 && SUB == 2) return launch_conv3x2_t<128, 2>(a0, a1, b, p, num_sms, st);
  if (BN == 128 && SUB == 1) return launch_conv3x2_t<128, 1>(a0, a1, b, p, num_sms, st);  // few tiles: finer N split
  if (BN == 64 && SUB == 2) return launch_conv3x2_t<64, 2>(a0, a1, b, p, num_sms, st);
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------
// CONV3R (conv3x3 forward / dgrad with 64 OUTPUT channels, "paired-tap" formulation, CTA pairs, M = 256).
//   The 64-column layers at full resolution (enc1.3, last.0, last.3 and the data gradients that produce 64 channels:
//   models/unet.py:53,66,69) are 20 % of the step's FLOPs.  In the tap-by-tap kernel above every one of the nine taps
//   re-reads its A window (128 rows x 32 B per MMA) against only 64 columns of tensor work: 180 KB of operand reads
//   per 128 x 64 output tile against 1152 clk of MMA, i.e. bound by the 128 B/clk shared-memory operand port
//   (measured: tensor pipe 32-38 %, shared-memory port 47-55 %), and it pays eleven barrier hand-shakes per tile.
//   Here two horizontal taps of a kernel row share ONE A window in ONE N = 128 MMA:
//       G0[h, w'] += X[h + r - 1, w'] * W[r, s=0]      (columns   0..63:  belongs to output column w' + 1)
//       G1[h, w'] += X[h + r - 1, w'] * W[r, s=1]      (columns 64..127:  belongs to output column w')
//   and the third tap is an N = 64 MMA on the window shifted by one pixel (+128 B: the swizzle is a function of the
//   absolute shared-memory address, scripts/desc_probe.cu), accumulating into G1:
//       G1[h, w'] += X[h + r - 1, w' + 1] * W[r, s=2]
//       out[h, w]  = G1[h, w] + G0[h, w - 1]
//   Per 128 x 64 tile: the same 1152 clk of MMA (3 rows x 4 k-steps x (64 + 32) clk), 132 KB of operand reads
//   (1031 clk of port time), 64 KB of TMEM reads (a three-column-group variant, N = 192, was measured first: its 96 KB
//   of accumulator reads per tile made the epilogue, not the tensor pipe, set the pace).  [W(r,0); W(r,1)] are adjacent
//   in the packed weights [tap = 3r + s][64][K], so they ARE the N = 128 B operand as they lie in memory.  The shift-and-
//   add G0[w - 1] happens in the epilogue: a warp owns one image-row segment of 32 pixels (its 32 TMEM lanes), so the
//   left neighbour comes by ONE warp shuffle; lanes 1..30 are outputs (lane 0 has no left neighbour, lane 31's shifted
//   window wraps into the next image row).  CTA r of the pair owns 4 image rows x 32 pixels (M = 128): halo tile 6 rows
//   x 32 px x 64 ch = 24 KB per 64-channel chunk, a plain K-major SWIZZLE_128B tile; kernel row r starts 32 rows (4 KB)
//   further down.  The whole weight set (K <= 128) stays resident in shared memory; four 128-column accumulator stages
//   decouple the MMA thread from the epilogue; BatchNorm statistics are accumulated per thread in registers across all
//   tiles and reduced across lanes ONCE per CTA.
constexpr int kRtTileW = 30;             // output pixels per row segment (lanes 1..30 of 32)
constexpr int kRtABytes = 6 * 32 * 128;  // halo tile of one 64-channel chunk
constexpr int kRtWBytes = 96 * 128;      // this CTA's share of one (chunk, kernel row) weight tile: 64 + 32 rows
constexpr int kRtAS = 4;                 // halo stages
constexpr int kRtOBytes = 4 * 30 * 128;  // output staging tile: 4 image rows x 30 pixels x 64 channels (TMA store)
constexpr int kRtAcc = 4;                // accumulator stages (128 TMEM columns each)
constexpr int kRtMaxKc = 2;              // 64-channel chunks of K kept resident
constexpr int kRtEpiThreads = 512;       // two groups of 8 epilogue warps, alternating tiles
constexpr int kRtThreads = 64 + kRtEpiThreads;
constexpr int kRtSmem = kRtAS * kRtABytes + kRtMaxKc * 3 * kRtWBytes + 2 * 16384 + 4 * 64 * 4 + (2 * kRtAS + 2 * kRtAcc + 1) * 8 + 16 + 1024;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kRtThreads, 1)
    igemm_conv3r_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                        const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapO,
                        const __grid_constant__ Conv3Params p) {
  constexpr int AS = kRtAS;
  constexpr int NACC = kRtAcc;
  constexpr uint32_t TMEM_COLS = 512;  // 4 accumulator stages x 128 columns

  if (d_pdl_mode == 0) pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sW = smem + AS * kRtABytes;
  uint8_t* sO = sW + kRtMaxKc * 3 * kRtWBytes;  // two 1024-aligned output staging tiles
  float* s_sum = reinterpret_cast<float*>(sO + 2 * 16384);
  float* s_sq = s_sum + 64;
  float* s_bias = s_sq + 64;
  float* s_aux = s_bias + 64;  // padding slot (keeps the barriers 8-byte aligned)
  uint64_t* a_full = reinterpret_cast<uint64_t*>(s_aux + 64);
  uint64_t* a_empty = a_full + AS;
  uint64_t* acc_full = a_empty + AS;
  uint64_t* acc_empty = acc_full + NACC;
  uint64_t* w_full = acc_empty + NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int nclusters = gridDim.x >> 1;
  const int kc = p.kc0 + p.kc1;
  const int total = p.m_tiles;  // one N tile (64 output channels)

  if (threadIdx.x == 0) {
    for (int s = 0; s < AS; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < NACC; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 16);
    }
    mbar_init(w_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
    tma_prefetch_desc(&mapO);
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  for (int i = threadIdx.x; i < 64; i += kRtThreads) {
    s_sum[i] = 0.f;
    s_sq[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (one per CTA)
    if (elect_one()) {
      // resident weights, per (chunk, kernel row): rows [0, 64) = this CTA's half of the N = 128 operand
      // [W(r,0); W(r,1)] (CTA 0 holds W(r,0), CTA 1 holds W(r,1)), rows [64, 96) = its half of W(r,2); three 32-row boxes
      if (leader) mbar_arrive_expect_tx(w_full, 2 * kc * 3 * kRtWBytes);
      for (int ch = 0; ch < kc; ++ch)
        for (int r = 0; r < 3; ++r) {
          uint8_t* dst = sW + (ch * 3 + r) * kRtWBytes;
          const int row0 = static_cast<int>(rank) * 64;
          tma_load_3d_2sm(dst, &mapB, leader_bar_addr(w_full), ch * 64, row0, r);
          tma_load_3d_2sm(dst + 32 * 128, &mapB, leader_bar_addr(w_full), ch * 64, row0 + 32, r);
          tma_load_3d_2sm(dst + 64 * 128, &mapB, leader_bar_addr(w_full), ch * 64, 128 + static_cast<int>(rank) * 32, r);
        }
      uint32_t ia = 0;
      for (int t = cluster_id; t < total; t += nclusters) {
        const int tx = t % p.tiles_w;
        const int rr = t / p.tiles_w;
        const int ty = rr % p.tiles_h;
        const int n = rr / p.tiles_h;
        const int w0 = tx * kRtTileW - 1;
        const int h0 = ty * 8 + 4 * static_cast<int>(rank) - 1;
        for (int ch = 0; ch < kc; ++ch, ++ia) {
          const uint32_t sa = ia % AS, pa = (ia / AS) & 1;
          mbar_wait(&a_empty[sa], pa ^ 1);
          if (leader) mbar_arrive_expect_tx(&a_full[sa], 2 * kRtABytes);
          if (ch < p.kc0)
            tma_load_5d_2sm(sA + sa * kRtABytes, &mapA0, leader_bar_addr(&a_full[sa]), ch * 64, w0, h0, n, 0);
          else
            tma_load_5d_2sm(sA + sa * kRtABytes, &mapA1, leader_bar_addr(&a_full[sa]), (ch - p.kc0) * 64, w0, h0, n, 0);
        }
      }
    }
    __syncwarp();
    if (d_pdl_mode == 1) pdl_launch_dependents();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer: leader CTA only
    if (leader && elect_one()) {
      constexpr uint32_t idesc128 = umma_idesc_bf16(256, 128, 0, 0);
      constexpr uint32_t idesc64 = umma_idesc_bf16(256, 64, 0, 0);
      mbar_wait(w_full, 0);
      tc_fence_after();
      const uint32_t wbase = smem_u32(sW);
      uint32_t ia = 0, it = 0;
      for (int t = cluster_id; t < total; t += nclusters, ++it) {
        const uint32_t acc = it % NACC, pacc = (it / NACC) & 1;
        mbar_wait(&acc_empty[acc], pacc ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * 128;
        for (int ch = 0; ch < kc; ++ch, ++ia) {
          const uint32_t sa = ia % AS, pa = (ia / AS) & 1;
          mbar_wait(&a_full[sa], pa);
          tc_fence_after();
          const uint32_t abase = smem_u32(sA + sa * kRtABytes);
          // descriptors differ from one MMA to the next only by a compile-time constant in the 14-bit start-address
          // field (all operand tiles lie below 256 KB, so the add never carries out of the field): ONE 64-bit add per
          // descriptor keeps the single issuing thread ahead of the tensor pipe (24 MMAs of 64 / 32 clk per chunk)
          const uint64_t adesc = umma_smem_desc(abase, 16, 1024);
          const uint64_t bdesc = umma_smem_desc(wbase + ch * 3 * kRtWBytes, 16, 1024);
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ao = static_cast<uint64_t>((r * 32 * 128 + k * 32) >> 4);
              const uint64_t bo = static_cast<uint64_t>((r * kRtWBytes + k * 32) >> 4);
              // taps s = 0 | s = 1 on the unshifted window -> G0 | G1
              umma_bf16_2cta(d0, adesc + ao, bdesc + bo, idesc128, (ch | r | k) != 0 ? 1u : 0u);
              // tap s = 2 on the window shifted by one pixel (one 128-byte row) -> G1 (always accumulates: the
              // N = 128 MMA in front of it has initialised the columns)
              umma_bf16_2cta(d0 + 64, adesc + ao + (128 >> 4), bdesc + bo + ((64 * 128) >> 4), idesc64, 1u);
            }
          }
          umma_commit_2cta(&a_empty[sa]);
        }
        umma_commit_2cta(&acc_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------ epilogue: 16 warps in two groups that take alternate tiles
    // (group g owns accumulator stages g and g + 2), so each warp has two tile periods for one tile; inside a group
    // quadrant q = warp % 4 = image row of this CTA's four, `half` = which 32 of the 64 output channels
    const int q = warp & 3;
    const int grp = (warp - 2) >> 3;
    const int half = ((warp - 2) >> 2) & 1;
    const int etid = threadIdx.x - 64;       // 0..511
    const int gtid = etid & 255;             // thread index inside the group
    const bool do_stats = p.stat_sum != nullptr;
    const bool affine = p.bn_scale != nullptr;
    for (int i = etid; i < 64; i += kRtEpiThreads) {
      s_bias[i] = (p.bias != nullptr && i < p.n_store) ? p.bias[i] : 0.f;
      if (affine) {  // inference: the statistics slots hold the BatchNorm scale / shift
        s_sum[i] = i < p.n_store ? p.bn_scale[i] : 0.f;
        s_sq[i] = i < p.n_store ? p.bn_shift[i] : 0.f;
      }
    }
    named_bar_sync(1, kRtEpiThreads);
    // BatchNorm statistics: thread e of a group sums one 16-byte chunk column (8 channels) of the STAGED bf16 tile over
    // the rows e / 8 + 32 i — 16 accumulators per thread instead of 64, and exactly the values that were stored
    float st_s[8], st_q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      st_s[j] = 0.f;
      st_q[j] = 0.f;
    }
    const int st_cc = gtid & 7, st_row0 = gtid >> 3;
    const float* bias_h = s_bias + half * 32;
    uint8_t* sbuf = sO + grp * 16384;
    uint32_t it = grp;
    for (int t = cluster_id + grp * nclusters; t < total; t += 2 * nclusters, it += 2) {
      const int tx = t % p.tiles_w;
      const int rr = t / p.tiles_w;
      const int ty = rr % p.tiles_h;
      const int n = rr / p.tiles_h;
      const uint32_t acc = it % NACC, pacc = (it / NACC) & 1;
      const int h0 = ty * 8 + 4 * static_cast<int>(rank);
      const int w0 = tx * kRtTileW;
      const bool out_lane = lane >= 1 && lane <= kRtTileW;
      mbar_wait(&acc_full[acc], pacc);
      tc_fence_after();
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 128 + half * 32;
      uint32_t g0[32], g1[32];
      tmem_ld32(tbase, g0);       // G0[l] belongs to the output one pixel to the right (lane l + 1)
      tmem_ld32(tbase + 64, g1);  // G1[l] belongs to this lane's output
      tmem_ld_wait();
      // the accumulator stage is in registers: hand it back to the MMA thread before the arithmetic
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&acc_empty[acc], 0);
#pragma unroll
      for (int j = 0; j < 32; ++j) g0[j] = __float_as_uint(__shfl_up_sync(0xffffffffu, __uint_as_float(g0[j]), 1));
      uint32_t pk[16];
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        float x[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = 2 * jj + e;
          float v = (__uint_as_float(g0[j]) + __uint_as_float(g1[j])) + bias_h[j];
          if (p.relu) v = fmaxf(v, 0.f);
          if (affine) v = fmaf(v, s_sum[half * 32 + j], s_sq[half * 32 + j]);
          x[e] = v;
        }
        pk[jj] = pack_bf16x2(x[0], x[1]);
      }
      // stage the tile in shared memory in the SWIZZLE_128B pattern of the store map (row = pixel, 16-byte chunk
      // index XOR row % 8: conflict-free), then ONE TMA store per CTA and tile writes full 128-byte lines and clips
      // at the image border; scattered 16-byte stores at a 128-byte stride were measured to cost ~1000 clk per tile.
      // The group's previous store (two tiles ago) has long finished reading this buffer: the wait below is free.
      if (gtid == 0) tma_store_wait_read<0>();
      named_bar_sync(2 + grp, 256);
      if (out_lane) {
        const int row = q * kRtTileW + lane - 1;
        uint8_t* rp = sbuf + row * 128;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          *reinterpret_cast<uint4*>(rp + (((half * 4 + jj) ^ (row & 7)) << 4)) =
              make_uint4(pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
      }
      fence_proxy_async();
      named_bar_sync(2 + grp, 256);
      if (gtid == 0) {
        tma_store_5d(&mapO, sbuf, 0, w0, h0, n, 0);
        tma_store_commit();
      }
      if (do_stats) {
        const bool interior = h0 + 4 <= p.H && w0 + kRtTileW <= p.W;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = st_row0 + 32 * i;
          bool ok = row < 4 * kRtTileW;
          if (ok && !interior) {
            const int qr = row / kRtTileW, px = row - qr * kRtTileW;
            ok = h0 + qr < p.H && w0 + px < p.W;
          }
          if (ok) {
            const uint4 v = *reinterpret_cast<const uint4*>(sbuf + row * 128 + ((st_cc ^ (row & 7)) << 4));
            const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const float lo = bf16lo_to_f32(u[jj]), hi = bf16hi_to_f32(u[jj]);
              st_s[2 * jj] += lo;
              st_s[2 * jj + 1] += hi;
              st_q[2 * jj] = fmaf(lo, lo, st_q[2 * jj]);
              st_q[2 * jj + 1] = fmaf(hi, hi, st_q[2 * jj + 1]);
            }
          }
        }
      }
    }
    if (gtid == 0) tma_store_wait<0>();
    if (do_stats) {
      // lanes l, l + 8, l + 16, l + 24 hold the same chunk column: fold them, then the 16 warps meet in shared memory,
      // then ONE fp64 atomic per channel and CTA
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        st_s[j] += __shfl_xor_sync(0xffffffffu, st_s[j], 8);
        st_s[j] += __shfl_xor_sync(0xffffffffu, st_s[j], 16);
        st_q[j] += __shfl_xor_sync(0xffffffffu, st_q[j], 8);
        st_q[j] += __shfl_xor_sync(0xffffffffu, st_q[j], 16);
      }
      if (lane < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          atomicAdd(&s_sum[lane * 8 + j], st_s[j]);
          atomicAdd(&s_sq[lane * 8 + j], st_q[j]);
        }
      }
      named_bar_sync(1, kRtEpiThreads);
      for (int i = etid; i < 64; i += kRtEpiThreads) {
        if (i < p.n_store) {
          atomicAdd(&p.stat_sum[i], static_cast<double>(s_sum[i]));
          atomicAdd(&p.stat_sq[i], static_cast<double>(s_sq[i]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
}

// a0/a1 box {64, 32, 6, 1, 1}; b = the packed weights viewed as [3][192][K], box {64, 32, 1}; p.tiles_w = ceil(W / 30),
// p.tiles_h = ceil(H / 8), p.m_tiles = N * tiles_h * tiles_w; kc0 + kc1 <= 2; o = the [N][H][W][64] destination, box
// {64, 30, 4, 1, 1}.
cudaError_t launch_conv3r(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& o,
                          const Conv3Params& p, int num_sms, cudaStream_t st) {
  if (p.kc0 + p.kc1 > kRtMaxKc || p.split_c != 0 || p.n_store > 64) return cudaErrorInvalidValue;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(igemm_conv3r_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRtSmem);
    if (e != cudaSuccess) return e;
  }
  int clusters = p.m_tiles;
  if (clusters > num_sms / 2) clusters = num_sms / 2;
  launch_k(igemm_conv3r_kernel, dim3(2 * clusters), dim3(kRtThreads), kRtSmem, st, a0, a1, b, o, p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// STEM (enc1.0, models/unet.py:50): conv3x3 of the 3-channel fp32 NCHW input, Cout = 64, as ONE kernel.  Round 1 wrote a
// K = 64 im2col matrix (134 MB at 16 x 256 x 256), read it back in a generic GEMM launch: ~7x the algorithmic bytes.
// Here 8 builder warps assemble the im2col tile [128 pixels][k = c*9 + r*3 + s, 27 of 32 used] directly in shared memory
// in the SWIZZLE_128B K-major layout the tensor core reads (from a 3 x 10 x 18 fp32 input patch staged per tile, zero
// outside the image = the conv padding), two K = 16 MMAs (M = 128, N = 64) multiply it with the resident [64][64]
// weight tile, and 4 epilogue warps do bias + ReLU (+ inference BatchNorm), stage the bf16 tile and write it with one
// TMA tensor store; BatchNorm statistics are summed from the staged tile.  Algorithmic bytes: 12 B in + 128 B out per
// pixel.  Tile = 8 image rows x 16 pixels (H % 8 == 0, W % 16 == 0: the U-Net needs multiples of 16 anyway).
constexpr int kStBuild = 128;                       // builder threads (warps 2..5): one im2col row (pixel) each
constexpr int kStEpi = 256;                         // epilogue threads (warps 6..13): (TMEM quadrant, 32-column half)
constexpr int kStThreads = 64 + kStBuild + kStEpi;  // + warp 0 (weights TMA), warp 1 (TMEM alloc + MMA issue)
constexpr int kStPatch = 3 * 10 * 18;               // fp32 input patch of one tile
constexpr int kStSmem = 2 * 16384 /*A*/ + 8192 /*W*/ + 2 * 16384 /*out staging*/ + 2 * kStPatch * 4 + 3 * 64 * 4 + 128 * 8 + 9 * 8 + 16 + 1024;

struct StemParams {
  const float* x;  // [N][Cin][H][W] fp32
  int N, Cin, H, W;
  int tiles_w, tiles_h, total;
  const float* bias;
  int relu;
  double* stat_sum;
  double* stat_sq;
  const float* bn_scale;
  const float* bn_shift;
};

__global__ void __launch_bounds__(kStThreads, 2)
    stem_conv_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapO,
                     const __grid_constant__ StemParams p) {
  if (d_pdl_mode == 0) pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sA = smem;                    // 2 x [128 px][128 B]
  uint8_t* sW = smem + 2 * 16384;        // [64 cout][128 B]
  uint8_t* sO = sW + 8192;               // 2 x [128 px][128 B]
  float* s_patch = reinterpret_cast<float*>(sO + 2 * 16384);  // 2 x [3][10][18]
  float* s_bias = s_patch + 2 * kStPatch;
  float* s_sum = s_bias + 64;
  float* s_sq = s_sum + 64;
  double* s_dstat = reinterpret_cast<double*>(s_sq + 64);  // [2][64] per-CTA statistics (fp64)
  uint64_t* a_full = reinterpret_cast<uint64_t*>(s_dstat + 128);
  uint64_t* a_empty = a_full + 2;
  uint64_t* acc_full = a_empty + 2;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool affine = p.bn_scale != nullptr;
  const bool do_stats = p.stat_sum != nullptr;
  for (int i = tid; i < 128; i += kStThreads) s_dstat[i] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], kStBuild / 32);
      mbar_init(&a_empty[s], 1);
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kStEpi / 32);
    }
    mbar_init(w_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&mapW);
    tma_prefetch_desc(&mapO);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);  // 2 accumulator stages x 64 columns
    tmem_relinquish();
  }
  // the unused half of every im2col row (k = 32..63) is zero for good; bias / inference-affine slots
  for (int i = tid; i < 256; i += kStThreads) {
    uint4* row = reinterpret_cast<uint4*>(sA + i * 128);
#pragma unroll
    for (int c = 4; c < 8; ++c) row[c ^ (i & 7)] = make_uint4(0u, 0u, 0u, 0u);
  }
  // predecessor grid complete and flushed from here on: the inference scale / shift are written by the kernel right
  // before this one (bn_finalize), so nothing that lives in global memory may be read above this line
  pdl_wait();
  for (int i = tid; i < 64; i += kStThreads) {
    s_bias[i] = p.bias != nullptr ? p.bias[i] : 0.f;
    s_sum[i] = affine ? p.bn_scale[i] : 0.f;
    s_sq[i] = affine ? p.bn_shift[i] : 0.f;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, 8192);
      tma_load_3d(sW, &mapW, w_full, 0, 0, 0);
    }
    __syncwarp();
    if (d_pdl_mode == 1) pdl_launch_dependents();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      mbar_wait(w_full, 0);
      tc_fence_after();
      const uint64_t bdesc = umma_smem_desc(smem_u32(sW), 16, 1024);
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.total; t += gridDim.x, ++it) {
        const uint32_t b = it & 1, ph = (it >> 1) & 1;
        mbar_wait(&acc_empty[b], ph ^ 1);
        mbar_wait(&a_full[b], ph);
        tc_fence_after();
        const uint64_t adesc = umma_smem_desc(smem_u32(sA + b * 16384), 16, 1024);
        umma_bf16(tmem_base + b * 64, adesc, bdesc, idesc, 0u);
        umma_bf16(tmem_base + b * 64, adesc + 2, bdesc + 2, idesc, 1u);  // k = 16..31 (+32 bytes)
        umma_commit(&a_empty[b]);
        umma_commit(&acc_full[b]);
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ------------------------------------------------ builders: im2col tile from the fp32 NCHW input, thread = pixel
    const int m = tid - 64;  // 0..127: pixel (m / 16, m % 16) of the tile
    const int pbase = (m >> 4) * 18 + (m & 15);
    // the (channel, row, column) of the up to five patch elements this thread stages per tile
    int pc[5], pr[5], pcol[5];
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      const int i = m + e * kStBuild;
      pc[e] = i / 180;
      pr[e] = (i - pc[e] * 180) / 18;
      pcol[e] = i - pc[e] * 180 - pr[e] * 18;
      if (i >= p.Cin * 180) pc[e] = -1;
    }
    const long long plane = static_cast<long long>(p.H) * p.W;
    // the patch values of a tile are requested one tile ahead (registers), so the global-load latency overlaps the
    // assembly of the previous tile
    auto load_patch = [&](int t, float (&pv)[5]) {
      const int tx = t % p.tiles_w;
      const int rr = t / p.tiles_w;
      const int ty = rr % p.tiles_h;
      const int n = rr / p.tiles_h;
      const int h0 = ty * 8 - 1, w0 = tx * 16 - 1;
      const float* xn = p.x + static_cast<long long>(n) * p.Cin * plane;
#pragma unroll
      for (int e = 0; e < 5; ++e) {
        const int hh = h0 + pr[e], ww = w0 + pcol[e];
        pv[e] = 0.f;
        if (pc[e] >= 0 && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) pv[e] = __ldg(xn + pc[e] * plane + hh * p.W + ww);
      }
    };
    float pv[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (static_cast<int>(blockIdx.x) < p.total) load_patch(blockIdx.x, pv);
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.total; t += gridDim.x, ++it) {
      const uint32_t b = it & 1, ph = (it >> 1) & 1;
      float* patch = s_patch + b * kStPatch;
      // (this patch buffer was last read two tiles ago, before the builders' previous barrier)
#pragma unroll
      for (int e = 0; e < 5; ++e)
        if (pc[e] >= 0) patch[m + e * kStBuild] = pv[e];
      named_bar_sync(4, kStBuild);
      if (t + static_cast<int>(gridDim.x) < p.total) load_patch(t + gridDim.x, pv);
      // im2col row of this pixel: k = c*9 + r*3 + s, 27 of 32 columns used (compile-time offsets into the patch)
      uint32_t pk[16];
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = 2 * jj + e;
          const int c = k / 9, rs = k - c * 9;
          const int r = rs / 3, s2 = rs - r * 3;
          v[e] = (k < 27 && c < p.Cin) ? patch[pbase + c * 180 + r * 18 + s2] : 0.f;
        }
        pk[jj] = pack_bf16x2(v[0], v[1]);
      }
      mbar_wait(&a_empty[b], ph ^ 1);  // the MMAs of tile it - 2 are done with this buffer
      uint4* row = reinterpret_cast<uint4*>(sA + b * 16384 + m * 128);
#pragma unroll
      for (int c = 0; c < 4; ++c) row[c ^ (m & 7)] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[b]);
    }
  } else {
    // ------------------------------------------------ epilogue: 8 warps = (TMEM lane quadrant q, 32-column half)
    const int q = warp & 3;
    const int half = (warp - 6) >> 2;
    const int etid = tid - 64 - kStBuild;  // 0..255
    const int m = q * 32 + lane;
    const int st_cc = etid & 7, st_row0 = etid >> 3;  // statistics: 16-byte chunk column, rows st_row0 + 32 i
    double st_s[8], st_q[8];  // fp64 totals: ~220 values per accumulator and CTA at 16 x 256 x 256
    float ts[8], tq[8];       // fp32 partials of up to 8 tiles
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      st_s[j] = 0.0;
      st_q[j] = 0.0;
      ts[j] = 0.f;
      tq[j] = 0.f;
    }
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.total; t += gridDim.x, ++it) {
      const uint32_t b = it & 1, ph = (it >> 1) & 1;
      const int tx = t % p.tiles_w;
      const int rr = t / p.tiles_w;
      const int ty = rr % p.tiles_h;
      const int n = rr / p.tiles_h;
      uint8_t* sbuf = sO + b * 16384;
      mbar_wait(&acc_full[b], ph);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * 64 + half * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);  // the accumulator stage is in registers
      uint32_t pk[16];
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        float x[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = half * 32 + 2 * jj + e;
          float f = __uint_as_float(v[2 * jj + e]) + s_bias[j];
          if (p.relu) f = fmaxf(f, 0.f);
          if (affine) f = fmaf(f, s_sum[j], s_sq[j]);
          x[e] = f;
        }
        pk[jj] = pack_bf16x2(x[0], x[1]);
      }
      if (etid == 0) tma_store_wait_read<1>();  // the store two tiles back has read this staging buffer
      named_bar_sync(5, kStEpi);
      uint8_t* rp = sbuf + m * 128;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(rp + (((half * 4 + c) ^ (m & 7)) << 4)) =
            make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      fence_proxy_async();
      named_bar_sync(5, kStEpi);
      if (etid == 0) {
        tma_store_5d(&mapO, sbuf, 0, tx * 16, ty * 8, n, 0);
        tma_store_commit();
      }
      if (do_stats) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = st_row0 + 32 * i;
          const uint4 u4 = *reinterpret_cast<const uint4*>(sbuf + row * 128 + ((st_cc ^ (row & 7)) << 4));
          const uint32_t u[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float lo = bf16lo_to_f32(u[jj]), hi = bf16hi_to_f32(u[jj]);
            ts[2 * jj] += lo;
            ts[2 * jj + 1] += hi;
            tq[2 * jj] = fmaf(lo, lo, tq[2 * jj]);
            tq[2 * jj + 1] = fmaf(hi, hi, tq[2 * jj + 1]);
          }
        }
        if ((it & 7) == 7) {  // fp32 partials over 8 tiles (32 values), then into the fp64 totals
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            st_s[j] += static_cast<double>(ts[j]);
            st_q[j] += static_cast<double>(tq[j]);
            ts[j] = 0.f;
            tq[j] = 0.f;
          }
        }
      }
    }
    if (etid == 0) tma_store_wait<0>();
    if (do_stats) {
      // lanes l, l + 8, l + 16, l + 24 hold the same chunk column; then ONE fp64 atomic per channel and warp
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        st_s[j] += static_cast<double>(ts[j]);
        st_q[j] += static_cast<double>(tq[j]);
        st_s[j] += __shfl_xor_sync(0xffffffffu, st_s[j], 8);
        st_s[j] += __shfl_xor_sync(0xffffffffu, st_s[j], 16);
        st_q[j] += __shfl_xor_sync(0xffffffffu, st_q[j], 8);
        st_q[j] += __shfl_xor_sync(0xffffffffu, st_q[j], 16);
      }
      // the 8 warps meet in shared memory, then ONE fp64 atomic per channel and CTA (2 x 64 hot addresses chip-wide)
      if (lane < 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          atomicAdd(&s_dstat[lane * 8 + j], st_s[j]);
          atomicAdd(&s_dstat[64 + lane * 8 + j], st_q[j]);
        }
      }
      named_bar_sync(5, kStEpi);
      if (etid < 64) atomicAdd(&p.stat_sum[etid], s_dstat[etid]);
      else if (etid < 128) atomicAdd(&p.stat_sq[etid - 64], s_dstat[etid]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

// w box {64, 64, 1} of the packed stem weights [64 cout][64 k]; o = NHWC map of y [N][H][W][64], box {64, 16, 8, 1, 1}
cudaError_t launch_stem_conv(const CUtensorMap& w, const CUtensorMap& o, const float* x, int N, int Cin, int H, int W,
                             const float* bias, int relu, double* stat_sum, double* stat_sq, const float* bn_scale,
                             const float* bn_shift, int num_sms, cudaStream_t st) {
  if (H % 8 || W % 16 || Cin * 9 > 32) return cudaErrorInvalidValue;
  StemParams p;
  p.x = x; p.N = N; p.Cin = Cin; p.H = H; p.W = W;
  p.tiles_w = W / 16; p.tiles_h = H / 8; p.total = N * p.tiles_h * p.tiles_w;
  p.bias = bias; p.relu = relu; p.stat_sum = stat_sum; p.stat_sq = stat_sq; p.bn_scale = bn_scale; p.bn_shift = bn_shift;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStSmem);
    if (e != cudaSuccess) return e;
  }
  int grid = p.total < 2 * num_sms ? p.total : 2 * num_sms;  // persistent: two 448-thread CTAs per SM (76 KB, <= 72 registers)
  launch_k(stem_conv_kernel, dim3(grid), dim3(kStThreads), kStSmem, st, w, o, p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// host side
template <int BN, int STAGES>
static constexpr int fprop_smem_bytes() {
  return STAGES * (128 * 128 + BN * 128) + kFpScratch + 3 * BN * 4 + (2 * STAGES + 4) * 8 + 16 + 1024;
}

template <int BN, int STAGES, typename OutT>
static cudaError_t launch_fprop_t(const CUtensorMap& a0, const CUtensorMap& a1,
                                  const CUtensorMap& b, const CUtensorMap& o, const FpropParams& p, int m_tiles,
                                  int n_tiles, cudaStream_t st) {
  constexpr int smem = fprop_smem_bytes<BN, STAGES>();
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(igemm_fprop_kernel<BN, STAGES, OutT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
  }
  int grid = m_tiles * n_tiles;
  const int cap = g_fprop_sms * (smem <= 113 * 1024 ? 2 : 1);  // resident CTAs per SM by shared memory
  if (grid > cap) grid = cap;
  launch_k(igemm_fprop_kernel<BN, STAGES, OutT>, dim3(grid), dim3(kFpThreads), smem, st, a0, a1, b, o, p, m_tiles, n_tiles);
  return cudaGetLastError();
}

cudaError_t launch_fprop(int BN, int out_is_f32, const CUtensorMap& a0, const CUtensorMap& a1,
                         const CUtensorMap& b, const FpropParams& p, int m_tiles, int n_tiles,
                         cudaStream_t st, const CUtensorMap* o) {
  if ((p.tma_store != 0) != (o != nullptr)) return cudaErrorInvalidValue;
  if (p.tma_store != 0 && (out_is_f32 || p.n_store % 64 != 0 || p.split_c != 0)) return cudaErrorInvalidValue;
  const CUtensorMap& om = o != nullptr ? *o : a0;  // unused unless p.tma_store
  if (out_is_f32) {
    if (BN == 32) return launch_fprop_t<32, 4, float>(a0, a1, b, om, p, m_tiles, n_tiles, st);
    return cudaErrorInvalidValue;
  }
  switch (BN) {
    case 64: return launch_fprop_t<64, 3, __nv_bfloat16>(a0, a1, b, om, p, m_tiles, n_tiles, st);
    case 128: return launch_fprop_t<128, 3, __nv_bfloat16>(a0, a1, b, om, p, m_tiles, n_tiles, st);
    case 256: return launch_fprop_t<256, 3, __nv_bfloat16>(a0, a1, b, om, p, m_tiles, n_tiles, st);
    default: return cudaErrorInvalidValue;
  }
}

template <int BN>
static cudaError_t launch_wgrad_t(const CUtensorMap& u, const CUtensorMap& t0,
                                  const CUtensorMap& t1, const WgradParams& p, cudaStream_t st) {
  const int stage_bytes = (2 + p.G * (BN / 64)) * kWgSlab;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > kWgStagesMax) stages = kWgStagesMax;
  if (stages < 2) return cudaErrorInvalidValue;
  const int smem = stages * stage_bytes + (2 * kWgStagesMax + 1) * 8 + 16 + 1024;
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(p.G * BN)) cols <<= 1;
  if (cols > 512) return cudaErrorInvalidValue;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(igemm_wgrad_kernel<BN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
  }
  dim3 grid(p.m_tiles * p.n_tiles * p.tap_groups, p.ksplit);
  launch_k(igemm_wgrad_kernel<BN>, dim3(grid), dim3(kThreads), smem, st, u, t0, t1, p, stages, cols);
  return cudaGetLastError();
}

cudaError_t launch_wgrad(int BN, const CUtensorMap& u, const CUtensorMap& t0, const CUtensorMap& t1,
                         const WgradParams& p, cudaStream_t st) {
  switch (BN) {
    case 64: return launch_wgrad_t<64>(u, t0, t1, p, st);
    case 128: return launch_wgrad_t<128>(u, t0, t1, p, st);
    default: return cudaErrorInvalidValue;
  }
}

// ------------------------------------------------------------------------------------------
// HEAD + LOSS + HEAD BACKWARD in one kernel (models/unet.py:72, trainer.py:113,174-175 for the head):
//   logits = z Wf^T + b        tcgen05, fp32 logits stay in tensor memory (never written to HBM)
//   CE (+ temperature-KL)      one thread per pixel on its TMEM row; dlogits -> bf16 -> shared memory (swizzled)
//   dz = dlogits Wd            tcgen05, A = the dlogits tile just written           -> bf16 [P][64]
//   dW += dlogits^T z          tcgen05, the SAME two tiles read MN-major, accumulated in TMEM over the CTA's tiles
//   db += colsum(dlogits)      per-thread sums, one warp-shuffle column sum per CTA
// HBM traffic: z read once, dz written once, labels (+ old logits) read once: 276 B/pixel instead of ~1.2 KB/pixel
// for the five separate launches.
// One persistent CTA per SM, 16 warps, software pipeline over 128-pixel tiles:
//   warp 0      TMA producer (z tiles, 4 stages)          warp 1   MMA issuer + TMEM owner
//   warps 4-7   loss group 0 (even tiles)                  warps 8-11  loss group 1 (odd tiles)
//   warps 12-15 dz store group (every tile)
// Tile i uses buffer b = i & 1 of the logits / dlogits / dz buffers, so loss group g always works on buffer g; the
// issuer runs logits(i+2) ahead of dz(i), and the loss of tile i+1 overlaps the dz store of tile i.
constexpr int kHlThreads = 512;
constexpr int kHlStages = 4;
constexpr int kHlTile = 128 * 128;                 // [128 px][64 ch] bf16
constexpr int kHlBars = 2 * kHlStages + 1 + 12;
constexpr int kHlSmem = (2 + kHlStages) * kHlTile + 32 * 128 + 64 * 128 + kHlBars * 8 + 16 + 1024;
constexpr uint32_t kHlTmemCols = 256;              // logits 2 x 32 | dz 2 x 64 | dW 64

// softmax cross-entropy (+ distillation) of one pixel: v = the 32 fp32 accumulators of its TMEM row.
// g[c] = gradient w.r.t. logit c (already scaled by gscale); same arithmetic as ce_kd_loss_kernel.
template <int NC, bool KD>
__device__ __forceinline__ void head_ce_row(const uint32_t (&v)[32], const float (&zo_in)[KD ? NC : 1], long long y,
                                            bool valid, const HeadLossParams& p, const float (&bias_r)[NC], float invT,
                                            float (&g)[NC], float& ce_local, float& kd_local) {
  const int C = p.C;
  float z[NC];
  float m4[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};  // 4 chains: instruction-level parallelism
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    z[c] = (c < C) ? __uint_as_float(v[c]) + bias_r[c] : -CUDART_INF_F;
    m4[c & 3] = fmaxf(m4[c & 3], z[c]);
  }
  const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
  float e[NC];
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    e[c] = __expf(z[c] - mx);  // exp(-inf) = 0 for the padding classes
    s4[c & 3] += e[c];
  }
  const float se = (s4[0] + s4[1]) + (s4[2] + s4[3]);
  const bool yok = y >= 0 && y < C;
  if (valid && !yok && p.err_flag != nullptr) *p.err_flag = 1;
  const bool live = valid && yok;
  const float inv_se = live ? p.gscale / se : 0.f;
  float zy = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    zy = (c == y) ? z[c] : zy;
    g[c] = e[c] * inv_se - ((c == y && live) ? p.gscale : 0.f);
  }
  if (live) ce_local += mx + __logf(se) - zy;
  if constexpr (KD) {
    // q = softmax(z[:Cold] / T), p0 = softmax(zold / T)
    float zo[NC];
    float mq = -CUDART_INF_F, mo = -CUDART_INF_F;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      z[c] = (c < p.Cold) ? z[c] * invT : -CUDART_INF_F;
      zo[c] = zo_in[c] * invT;  // -inf beyond Cold
      mq = fmaxf(mq, z[c]);
      mo = fmaxf(mo, zo[c]);
    }
    float sq = 0.f, so = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      sq += __expf(z[c] - mq);
      so += __expf(zo[c] - mo);
    }
    const float lq = mq + __logf(sq), lo = mo + __logf(so);
    const float kscale = valid ? p.lambda * p.T * p.gscale : 0.f;
    float kd = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c < p.Cold) {
        const float logq = z[c] - lq, logp0 = zo[c] - lo;
        const float p0 = __expf(logp0);
        kd += p0 * (logp0 - logq);
        g[c] += kscale * (__expf(logq) - p0);
      }
    }
    if (valid) kd_local += kd;
  }
}

// NC: compile-time bound on the class count (21 = the reference head, 32 = generic); KD: distillation term present
template <int NC, bool KD>
__global__ void __launch_bounds__(kHlThreads, 1)
    head_loss_kernel(const __grid_constant__ CUtensorMap mapZ, const __grid_constant__ CUtensorMap mapWf,
                     const __grid_constant__ CUtensorMap mapWd, const __grid_constant__ HeadLossParams p) {
  if (d_pdl_mode == 0) pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sD = smem;                          // 2 x dlogits tile [128 px][64 cls] bf16, SWIZZLE_128B rows
  uint8_t* sZ = smem + 2 * kHlTile;            // z stages (whatever follows a dlogits tile is the ignored upper half of M in dW)
  uint8_t* sWf = sZ + kHlStages * kHlTile;     // [32 cls][64 ch]
  uint8_t* sWd = sWf + 32 * 128;               // [64 ch][64 cls]
  uint64_t* z_full = reinterpret_cast<uint64_t*>(sWd + 64 * 128);
  uint64_t* z_empty = z_full + kHlStages;
  uint64_t* w_full = z_empty + kHlStages;
  uint64_t* lg_full = w_full + 1;    // [2] logits ready           (tcgen05.commit)
  uint64_t* lg_empty = lg_full + 2;  // [2] logits read out         (4 warps)
  uint64_t* d_full = lg_empty + 2;   // [2] dlogits tile written    (4 warps)
  uint64_t* d_empty = d_full + 2;    // [2] dlogits tile consumed   (tcgen05.commit)
  uint64_t* dz_full = d_empty + 2;   // [2] dz ready                (tcgen05.commit)
  uint64_t* dz_empty = dz_full + 2;  // [2] dz read out             (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dz_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long tiles = (p.P + 127) / 128;
  const uint32_t n_local = blockIdx.x < tiles ? static_cast<uint32_t>((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
  if (tid == 0) {
    for (int s = 0; s < kHlStages; ++s) {
      mbar_init(&z_full[s], 1);
      mbar_init(&z_empty[s], 1);
    }
    mbar_init(w_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&lg_full[b], 1);
      mbar_init(&lg_empty[b], 4);
      mbar_init(&d_full[b], 4);
      mbar_init(&d_empty[b], 1);
      mbar_init(&dz_full[b], 1);
      mbar_init(&dz_empty[b], 4);
    }
    fence_mbar_init();
    tma_prefetch_desc(&mapZ);
    tma_prefetch_desc(&mapWf);
    tma_prefetch_desc(&mapWd);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kHlTmemCols);
    tmem_relinquish();
  }
  if (tid < 256) {  // the zero half of every dlogits row (classes 32..63) is written once
    uint4* row = reinterpret_cast<uint4*>(sD + tid * 128);
#pragma unroll
    for (int c = 4; c < 8; ++c) row[c ^ (tid & 7)] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, 32 * 128 + 64 * 128);
      tma_load_3d(sWf, &mapWf, w_full, 0, 0, 0);
      tma_load_3d(sWd, &mapWd, w_full, 0, 0, 0);
      for (uint32_t i = 0; i < n_local; ++i) {
        const uint32_t s = i % kHlStages, k = i / kHlStages;
        mbar_wait(&z_empty[s], (k & 1) ^ 1);
        mbar_arrive_expect_tx(&z_full[s], kHlTile);
        const long long t = blockIdx.x + static_cast<long long>(i) * gridDim.x;
        tma_load_5d(sZ + s * kHlTile, &mapZ, &z_full[s], 0, static_cast<int>(t * 128), 0, 0, 0);
      }
    }
    __syncwarp();
    if (d_pdl_mode == 1) pdl_launch_dependents();
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (elect_one() && n_local > 0) {
      constexpr uint32_t idesc_logits = umma_idesc_bf16(128, 32, 0, 0);
      constexpr uint32_t idesc_dz = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_dw = umma_idesc_bf16(128, 64, 1, 1);
      const uint32_t wf = smem_u32(sWf), wd = smem_u32(sWd);
      mbar_wait(w_full, 0);
      auto logits = [&](uint32_t i) {  // logits[128][32] = Z[128][64] Wf^T
        const uint32_t b = i & 1, k = i >> 1, s = i % kHlStages;
        mbar_wait(&z_full[s], (i / kHlStages) & 1);
        mbar_wait(&lg_empty[b], (k & 1) ^ 1);
        tc_fence_after();
        const uint32_t zaddr = smem_u32(sZ + s * kHlTile);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16(tmem_base + 32 * b, umma_smem_desc(zaddr + kk * 32, 16, 1024), umma_smem_desc(wf + kk * 32, 16, 1024),
                    idesc_logits, kk != 0 ? 1u : 0u);
        umma_commit(&lg_full[b]);
      };
      logits(0);
      if (n_local > 1) logits(1);
      for (uint32_t i = 0; i < n_local; ++i) {
        // logits of tile i + 2 first: its buffer is free as soon as the loss group has READ the logits of tile i, so the
        // MMA round trip overlaps the loss arithmetic of tile i instead of sitting between two tiles of that group
        if (i + 2 < n_local) logits(i + 2);
        const uint32_t b = i & 1, k = i >> 1, s = i % kHlStages;
        mbar_wait(&d_full[b], k & 1);
        mbar_wait(&dz_empty[b], (k & 1) ^ 1);
        tc_fence_after();
        const uint32_t daddr = smem_u32(sD + b * kHlTile);
        const uint32_t zaddr = smem_u32(sZ + s * kHlTile);
        // dz[128][64] = D[128][32] Wd^T
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
          umma_bf16(tmem_base + 64 + 64 * b, umma_smem_desc(daddr + kk * 32, 16, 1024),
                    umma_smem_desc(wd + kk * 32, 16, 1024), idesc_dz, kk != 0 ? 1u : 0u);
        // dW[cls][ch] += D^T Z through MN-major views of the same two tiles: A = D^T (M = class; rows 64..127 of M
        // read the 16 KB after the tile and are never used), B = Z (N = channel), K = 128 pixels in 8 steps of 16 rows
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16(tmem_base + 192, umma_smem_desc(daddr + kk * 2048, kHlTile, 1024),
                    umma_smem_desc(zaddr + kk * 2048, kHlTile, 1024), idesc_dw, (i | kk) != 0 ? 1u : 0u);
        umma_commit(&dz_full[b]);
        umma_commit(&d_empty[b]);
        umma_commit(&z_empty[s]);
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 12) {
    // ------------------------------------------------ loss groups: thread = pixel = TMEM lane
    const uint32_t grp = (warp - 4) >> 2;  // = buffer index
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 32 * grp;
    uint8_t* dtile = sD + grp * kHlTile;
    const float invT = 1.f / p.T;
    float ce_local = 0.f, kd_local = 0.f;
    float g[NC], db_acc[NC], bias_r[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      db_acc[c] = 0.f;
      bias_r[c] = (c < p.C && p.bias != nullptr) ? p.bias[c] : 0.f;
    }
    for (uint32_t i = grp; i < n_local; i += 2) {
      const uint32_t k = i >> 1;
      const long long t = blockIdx.x + static_cast<long long>(i) * gridDim.x;
      const long long pix = t * 128 + row;
      const bool valid = pix < p.P;
      // the label (and the old model's logits) are requested before waiting for the tensor core
      const long long y = valid ? p.labels[pix] : 0;
      float zo[KD ? NC : 1];
      if constexpr (KD) {
        const float* orow = p.old_logits + (valid ? pix : 0) * p.Cold;
#pragma unroll
        for (int c = 0; c < NC; ++c) zo[c] = (c < p.Cold) ? orow[c] : -CUDART_INF_F;
      }
      mbar_wait(&lg_full[grp], k & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(lane_addr, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&lg_empty[grp]);  // the issuer may compute the logits of tile i + 2
      head_ce_row<NC, KD>(v, zo, y, valid, p, bias_r, invT, g, ce_local, kd_local);
      // dlogits row -> bf16 -> shared memory in the SWIZZLE_128B pattern TMA would have produced
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        pk[j] = (2 * j < NC) ? pack_bf16x2(g[2 * j], (2 * j + 1 < NC) ? g[2 * j + 1] : 0.f) : 0u;
      mbar_wait(&d_empty[grp], (k & 1) ^ 1);  // the MMAs of tile i - 2 are done with this buffer
      uint4* drow = reinterpret_cast<uint4*>(dtile + row * 128);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        drow[c ^ (row & 7)] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&d_full[grp]);
#pragma unroll
      for (int j = 0; j < 16; ++j) {  // the bias gradient sums what the tensor core sees (bf16-rounded)
        if (2 * j < NC) db_acc[2 * j] += bf16lo_to_f32(pk[j]);
        if (2 * j + 1 < NC) db_acc[2 * j + 1] += bf16hi_to_f32(pk[j]);
      }
    }
    {  // bias gradient: per-thread sums over this CTA's tiles -> one column sum per warp
      float f[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) f[c] = (c < NC) ? db_acc[c] : 0.f;
      const float colsum = warp_colsum32(f, lane);
      if (lane < p.C) atomicAdd(&p.dbias[lane], static_cast<double>(colsum));
    }
    for (int o = 16; o > 0; o >>= 1) {
      ce_local += __shfl_xor_sync(0xffffffffu, ce_local, o);
      kd_local += __shfl_xor_sync(0xffffffffu, kd_local, o);
    }
    if (lane == 0) {
      atomicAdd(p.loss_acc, static_cast<double>(ce_local));
      if (KD) atomicAdd(p.loss_acc + 1, static_cast<double>(kd_local));
    }
  } else if (warp >= 12) {
    // ------------------------------------------------ dz store group
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (uint32_t i = 0; i < n_local; ++i) {
      const uint32_t b = i & 1, k = i >> 1;
      const long long t = blockIdx.x + static_cast<long long>(i) * gridDim.x;
      const long long pix = t * 128 + row;
      const bool valid = pix < p.P;
      mbar_wait(&dz_full[b], k & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(lane_base + 64 + 64 * b, v0);
      tmem_ld32(lane_base + 64 + 64 * b + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dz_empty[b]);
      if (valid) {
        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.dz) + pix * 64);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          o[j] = make_uint4(pack_bf16x2(__uint_as_float(v0[8 * j]), __uint_as_float(v0[8 * j + 1])),
                            pack_bf16x2(__uint_as_float(v0[8 * j + 2]), __uint_as_float(v0[8 * j + 3])),
                            pack_bf16x2(__uint_as_float(v0[8 * j + 4]), __uint_as_float(v0[8 * j + 5])),
                            pack_bf16x2(__uint_as_float(v0[8 * j + 6]), __uint_as_float(v0[8 * j + 7])));
#pragma unroll
        for (int j = 0; j < 4; ++j)
          o[4 + j] = make_uint4(pack_bf16x2(__uint_as_float(v1[8 * j]), __uint_as_float(v1[8 * j + 1])),
                                pack_bf16x2(__uint_as_float(v1[8 * j + 2]), __uint_as_float(v1[8 * j + 3])),
                                pack_bf16x2(__uint_as_float(v1[8 * j + 4]), __uint_as_float(v1[8 * j + 5])),
                                pack_bf16x2(__uint_as_float(v1[8 * j + 6]), __uint_as_float(v1[8 * j + 7])));
      }
    }
    // the last dz_full also covers the last dW MMA: flush the weight gradient (TMEM lanes 0..31 = classes)
    if (q == 0 && n_local > 0) {
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(tmem_base + 192 + half * 32, v);
        tmem_ld_wait();
        if (lane < p.C) {
          float* o = p.dw + lane * 64 + half * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            red_add_v4(o + 4 * j, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                       __uint_as_float(v[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kHlTmemCols);
}

template <int NC, bool KD>
static cudaError_t launch_head_loss_t(const CUtensorMap& z, const CUtensorMap& wf, const CUtensorMap& wd,
                                      const HeadLossParams& p, int num_sms, cudaStream_t st) {
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(head_loss_kernel<NC, KD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHlSmem);
    if (e != cudaSuccess) return e;
  }
  long long grid = (p.P + 127) / 128;
  if (grid > num_sms) grid = num_sms;
  launch_k(head_loss_kernel<NC, KD>, dim3(static_cast<unsigned>(grid)), dim3(kHlThreads), kHlSmem, st, z, wf, wd, p);
  return cudaGetLastError();
}

cudaError_t launch_head_loss(const CUtensorMap& z, const CUtensorMap& wf, const CUtensorMap& wd, const HeadLossParams& p,
                             int num_sms, cudaStream_t st) {
  const bool kd = p.old_logits != nullptr;
  if (p.C == 21) {  // the reference head (21 VOC classes): loops unrolled to exactly 21 columns
    return kd ? launch_head_loss_t<21, true>(z, wf, wd, p, num_sms, st)
              : launch_head_loss_t<21, false>(z, wf, wd, p, num_sms, st);
  }
  return kd ? launch_head_loss_t<32, true>(z, wf, wd, p, num_sms, st)
            : launch_head_loss_t<32, false>(z, wf, wd, p, num_sms, st);
}

// ------------------------------------------------------------------------------------------
// HEAD + ARGMAX + CONFUSION MATRIX in one kernel: the statistics / validation path of the reference
// (models/unet.py:72 -> argmax(softmax(out), 1), .eq(labels).sum(), metrics._fast_conf_matrix: trainer.py:183-184,
// 279, metrics.py:32-38).  The logits stay in tensor memory: z is read once, nothing but the optional prediction map
// is written.  128 threads per CTA (thread = pixel = TMEM lane), two logits buffers so that the MMA of tile i + 1
// runs under the arg-max of tile i, several CTAs per SM; per-warp shared-memory histograms, int64 global bins.
// Same tie / NaN rule as argmax_confusion_kernel (first maximum wins, a NaN beats every number).
constexpr int kHaThreads = 128;
__global__ void __launch_bounds__(kHaThreads, 4)
    head_argmax_kernel(const __grid_constant__ CUtensorMap mapZ, const __grid_constant__ CUtensorMap mapWf,
                       const __grid_constant__ HeadArgmaxParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sZ = smem;                       // two z stages
  uint8_t* sWf = sZ + 2 * kHlTile;          // [32 cls][64 ch]
  uint64_t* z_full = reinterpret_cast<uint64_t*>(sWf + 32 * 128);
  uint64_t* w_full = z_full + 2;
  uint64_t* lg_full = w_full + 1;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lg_full + 2);
  unsigned int* hist = reinterpret_cast<unsigned int*>(tmem_slot + 4);  // [4 warps][nc * nc]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nbins = p.nc * p.nc;
  const long long tiles = (p.P + 127) / 128;
  const uint32_t n_local = blockIdx.x < tiles ? static_cast<uint32_t>((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
  if (tid == 0) {
    mbar_init(&z_full[0], 1);
    mbar_init(&z_full[1], 1);
    mbar_init(w_full, 1);
    mbar_init(&lg_full[0], 1);
    mbar_init(&lg_full[1], 1);
    fence_mbar_init();
    tma_prefetch_desc(&mapZ);
    tma_prefetch_desc(&mapWf);
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  if (p.conf != nullptr)
    for (int i = tid; i < 4 * nbins; i += kHaThreads) hist[i] = 0u;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  constexpr uint32_t idesc = umma_idesc_bf16(128, 32, 0, 0);
  auto tile_of = [&](uint32_t i) { return blockIdx.x + static_cast<long long>(i) * gridDim.x; };
  auto issue_logits = [&](uint32_t i) {  // thread 0: logits[128][32] of local tile i into buffer i & 1
    const uint32_t s = i & 1;
    mbar_wait(&z_full[s], (i >> 1) & 1);
    tc_fence_after();
    const uint32_t zaddr = smem_u32(sZ + s * kHlTile);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16(tmem_base + 32 * s, umma_smem_desc(zaddr + k * 32, 16, 1024),
                umma_smem_desc(smem_u32(sWf) + k * 32, 16, 1024), idesc, k != 0 ? 1u : 0u);
    umma_commit(&lg_full[s]);
  };
  if (tid == 0 && n_local > 0) {
    mbar_arrive_expect_tx(w_full, 32 * 128);
    tma_load_3d(sWf, &mapWf, w_full, 0, 0, 0);
    for (uint32_t i = 0; i < 2 && i < n_local; ++i) {
      mbar_arrive_expect_tx(&z_full[i], kHlTile);
      tma_load_5d(sZ + i * kHlTile, &mapZ, &z_full[i], 0, static_cast<int>(tile_of(i) * 128), 0, 0, 0);
    }
    mbar_wait(w_full, 0);
    issue_logits(0);
  }

  float bias_r[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) bias_r[c] = (c < p.C && p.bias != nullptr) ? p.bias[c] : 0.f;
  unsigned int* mine = hist + warp * nbins;
  const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  unsigned int ok = 0;
  for (uint32_t i = 0; i < n_local; ++i) {
    const uint32_t s = i & 1;
    if (tid == 0 && i + 1 < n_local) issue_logits(i + 1);  // its buffer was drained before the barrier below (tile i - 1)
    const long long pix = tile_of(i) * 128 + tid;
    const bool valid = pix < p.P;
    const long long t = valid ? p.labels[pix] : -1;
    mbar_wait(&lg_full[s], (i >> 1) & 1);
    tc_fence_after();
    uint32_t v[32];
    tmem_ld32(lane_addr + 32 * s, v);
    tmem_ld_wait();
    float best = __uint_as_float(v[0]) + bias_r[0];
    int bi = 0;
#pragma unroll
    for (int c = 1; c < 32; ++c) {
      const float x = __uint_as_float(v[c]) + bias_r[c];
      if (c < p.C && (x > best || (x != x && best == best))) {
        best = x;
        bi = c;
      }
    }
    if (valid) {
      if (p.pred_out != nullptr) p.pred_out[pix] = bi;
      if (t == bi) ++ok;
      if (p.conf != nullptr && t >= 0 && t < p.nc) atomicAdd(mine + t * p.nc + bi, 1u);
    }
    tc_fence_before();
    __syncthreads();  // logits buffer s and (its MMA being complete) z stage s are free
    if (tid == 0 && i + 2 < n_local) {
      mbar_arrive_expect_tx(&z_full[s], kHlTile);
      tma_load_5d(sZ + s * kHlTile, &mapZ, &z_full[s], 0, static_cast<int>(tile_of(i + 2) * 128), 0, 0, 0);
    }
  }
  for (int o = 16; o > 0; o >>= 1) ok += __shfl_xor_sync(0xffffffffu, ok, o);
  if (lane == 0 && ok && p.correct != nullptr) atomicAdd(p.correct, static_cast<unsigned long long>(ok));
  __syncthreads();
  if (p.conf != nullptr) {
    for (int b = tid; b < nbins; b += kHaThreads) {
      unsigned long long sum = 0;
      for (int w = 0; w < 4; ++w) sum += hist[w * nbins + b];
      if (sum) atomicAdd(p.conf + b, sum);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

cudaError_t launch_head_argmax(const CUtensorMap& z, const CUtensorMap& wf, const HeadArgmaxParams& p, int num_sms,
                               cudaStream_t st) {
  const int smem = 2 * kHlTile + 32 * 128 + 5 * 8 + 16 + 4 * p.nc * p.nc * 4 + 1024;
  if (smem > 100 * 1024) return cudaErrorInvalidValue;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(head_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return e;
  }
  long long grid = (p.P + 127) / 128;
  if (grid > 4ll * num_sms) grid = 4ll * num_sms;
  launch_k(head_argmax_kernel, dim3(static_cast<unsigned>(grid)), dim3(kHaThreads), smem, st, z, wf, p);
  return cudaGetLastError();
}

}  // namespace clk
