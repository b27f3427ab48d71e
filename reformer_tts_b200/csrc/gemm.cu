// bf16 GEMM with fused epilogue on tcgen05 (rtts_gemm_bf16): projections (rp R1), FeedForward
// (ref:reformer_tts/model/modules.py:195-207) and their backward (dgrad / wgrad, split-K atomic).
//
// C[M,N] = epilogue(A . B^T), 128x128 output tile per CTA, K consumed in 64-wide blocks through a
// 3-stage TMA -> shared-memory ring (SWIZZLE_128B), accumulator in TMEM (128 fp32 columns), so two
// CTAs fit per SM and one CTA's epilogue overlaps the other's main loop.
//   warp 0    : TMA producer (one elected lane)
//   warp 1    : tcgen05.mma issuer (one elected lane)
//   warps 2-9 : epilogue, two warps per TMEM lane quarter (each half of the tile's columns): tcgen05.ld -> bias / ReLU / gate / cast /
//               dropout / residual -> global
// Operands may be K-major (row-major [rows, K]) or MN-major (stored [K, rows]); the latter is what the
// weight-gradient GEMM dW = dY^T X needs, and is loaded as two 64-wide boxes per 128-row tile.
#include <cuda.h>

#include "common.cuh"
#include "host_util.h"
#include "rtts_b200.h"
#include "tma_host.h"

namespace rtts {

constexpr int kBM = 128, kBK = 64;
// Epilogue warps: two per TMEM lane quarter, each taking half of the tile's columns.  The K = 512 projections (8 k-blocks per tile) are
// bound by the epilogue, not by the main loop: 20480x512x512 went 33 -> 26 us, the training step 24.2 -> 23.3 ms.  (Four warps for the
// 256-column tiles only was measured too: 24.2 ms.)
__host__ __device__ constexpr int gemm_epi_warps(int bn) { return bn > 0 ? 8 : 4; }
__host__ __device__ constexpr int gemm_threads(int bn) { return 64 + gemm_epi_warps(bn) * 32; }
constexpr int kATileBytes = kBM * kBK * 2;                    // 16 KB

template <int BN>
struct GemmCfg {
  static constexpr int kBTileBytes = BN * kBK * 2;            // 16 | 32 KB
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kStages = BN == 128 ? 6 : 4;           // 192 KB of operand ring either way
  static constexpr int kOffStaging = kStages * kStageBytes + 256;                 // after the barriers: 8 epilogue warps x one 32-row x 128-B tile (gate rows, then output rows)
  static constexpr int kSmem = kOffStaging + 8 * 4096 + 1024 /*align*/;
  static constexpr uint32_t kTmemCols = 2 * BN;               // two accumulators: the epilogue of tile i overlaps the main loop of tile i+1
};

struct GemmParams {
  void* C;
  int64_t ldc;
  const float* bias;
  const __nv_bfloat16* gate;
  int64_t ldgate;
  float* colsum;
  const uint8_t* keep;   // dropout keep mask [M,N] (1 = keep), row stride ldkeep; nullptr = no dropout
  int64_t ldkeep;
  float keep_scale;      // 1 / (1 - p)
  int M, N, K;       // K = per-split extent
  int epilogue;
  int tiles_n, tiles_mn, work;   // work = tiles_mn * split_k
};

// Sum r[0..32) over the 32 lanes of a warp; lane i ends up owning column i's total.
__device__ __forceinline__ float warp_column_sums(float* r, int lane) {
#pragma unroll
  for (int width = 16; width >= 1; width >>= 1) {
    const bool upper = (lane & width) != 0;
#pragma unroll
    for (int i = 0; i < width; ++i) {
      const float send = upper ? r[i] : r[i + width];
      const float keep = upper ? r[i + width] : r[i];
      r[i] = keep + __shfl_xor_sync(0xffffffffu, send, width);
    }
  }
  return r[0];
}

// Persistent: one CTA per SM walks work items w = blockIdx.x, += gridDim.x; w -> (split, m tile, n tile), n fastest so that
// CTAs running at the same time share A panels in L2.
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(gemm_threads(BN), 1) gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                    const __grid_constant__ CUtensorMap tmap_b,
                                                                    const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t s_base = smem_u32(smem);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* acc_full = empty_bar + kStages;     // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = p.K / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full + a, 1);
      mbar_init(acc_empty + a, gemm_epi_warps(BN) * 32);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;       // running k-block counter across work items: the operand ring never drains between tiles
      for (int w = blockIdx.x; w < p.work; w += gridDim.x) {
        const int split = w / p.tiles_mn, tile = w - split * p.tiles_mn;
        const int m0 = (tile / p.tiles_n) * kBM, n0 = (tile % p.tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(empty_bar + s, ((it / kStages) & 1) ^ 1);
          const uint32_t sa = s_base + s * Cfg::kStageBytes, sb = sa + kATileBytes;
          const int k0 = split * p.K + kb * kBK;
          mbar_arrive_expect_tx(full_bar + s, Cfg::kStageBytes);
          if (A_MN) {
            tma_load_2d(sa, &tmap_a, full_bar + s, m0, k0);
            tma_load_2d(sa + kATileBytes / 2, &tmap_a, full_bar + s, m0 + 64, k0);
          } else {
            tma_load_2d(sa, &tmap_a, full_bar + s, k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int nb = 0; nb < BN / 64; ++nb) tma_load_2d(sb + nb * (64 * kBK * 2), &tmap_b, full_bar + s, n0 + nb * 64, k0);
          } else {
#pragma unroll
            for (int nb = 0; nb < BN / 128; ++nb) tma_load_2d(sb + nb * (128 * kBK * 2), &tmap_b, full_bar + s, k0, n0 + nb * 128);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {      // elect.sync: straight-line tcgen05.mma issue (a plain lane test makes the compiler wrap every MMA in a uniformity loop)
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN, A_MN, B_MN);
      constexpr uint32_t hi = umma_desc_hi_sw128(1024);
      int it = 0, t = 0;
      for (int w = blockIdx.x; w < p.work; w += gridDim.x, ++t) {
        const int acc = t & 1;
        mbar_wait(acc_empty + acc, ((t >> 1) & 1) ^ 1);      // epilogue has drained this accumulator (two tiles ago)
        tc_fence_after_sync();
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(full_bar + s, (it / kStages) & 1);
          tc_fence_after_sync();
          const uint32_t sa = s_base + s * Cfg::kStageBytes, sb = sa + kATileBytes;
          const uint32_t da0 = A_MN ? umma_desc_lo(sa, kATileBytes / 2) : umma_desc_lo(sa, 16);
          const uint32_t db0 = B_MN ? umma_desc_lo(sb, 64 * kBK * 2) : umma_desc_lo(sb, 16);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_ss_lo(tmem + acc * BN, da0 + ((A_MN ? k * 2048 : k * 32) >> 4), db0 + ((B_MN ? k * 2048 : k * 32) >> 4), hi, idesc, (kb | k) != 0);
          umma_commit(empty_bar + s);          // smem slot reusable once these MMAs retire
        }
        umma_commit(acc_full + acc);           // accumulator complete
      }
    }
  } else {
    // ---- epilogue: warp w owns TMEM lanes [32*(w%4), +32) = output rows m0 + that range ----
    const int quarter = warp & 3;
    const int epi = p.epilogue;
    // Staging tile of this warp (32 rows x 128 B, 16-byte chunks XOR-swizzled by the row): the thread-per-row register layout that
    // tcgen05.ld produces is transposed through it so that global accesses are whole 128-byte row segments (four rows per warp
    // instruction) instead of 32 rows x 16 B.
    // K = 512 GEMMs (8 k-blocks per tile) are bound by this epilogue, not by the main loop: two warps per lane quarter split the columns
    const int chalf = (warp - 2) >> 2;
    const uint32_t stage = s_base + Cfg::kOffStaging + (warp - 2) * 4096, gstage = stage;      // one tile: gate rows first, then output rows
    const int srow = lane >> 3, sch = lane & 7;      // coalesced phase: lane -> (row within a group of four, 16-byte chunk)
    int t = 0;
    for (int w = blockIdx.x; w < p.work; w += gridDim.x, ++t) {
      const int split = w / p.tiles_mn, tile = w - split * p.tiles_mn;
      const int m0 = (tile / p.tiles_n) * kBM, n0 = (tile % p.tiles_n) * BN;
      const int acc = t & 1;
      const int row = m0 + quarter * 32 + lane;
      // residual rows / keep-mask bytes of this tile: pulled into L2 while the main loop of the tile is still running, so that the
      // epilogue's loads (issued per 32-column chunk) do not each pay an HBM round trip
      if (epi & (RTTS_EPI_RESID_ADD | RTTS_EPI_RESID_SUB)) {
        const char* rp = reinterpret_cast<const char*>(reinterpret_cast<const float*>(p.gate) + static_cast<int64_t>(row) * p.ldgate + n0);
#pragma unroll
        for (int i = 0; i < BN * 4 / 128; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + i * 128));
      }
      if (p.keep != nullptr) {
        const char* kp = reinterpret_cast<const char*>(p.keep + static_cast<int64_t>(row) * p.ldkeep + n0);
#pragma unroll
        for (int i = 0; i < (BN + 127) / 128; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(kp + i * 128));
      }
      mbar_wait(acc_full + acc, (t >> 1) & 1);
      tc_fence_after_sync();
#pragma unroll 1
      constexpr int kColsPerWarp = BN / (gemm_epi_warps(BN) / 4);
      for (int c0 = chalf * kColsPerWarp; c0 < (chalf + 1) * kColsPerWarp; c0 += 32) {
        uint32_t raw[32];
        tmem_ld32(tmem + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + c0, raw);
        tmem_ld_wait();
        float r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __uint_as_float(raw[i]);
        const int col = n0 + c0;
        if (epi & RTTS_EPI_BIAS) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + col + i));
            r[i] += bv.x; r[i + 1] += bv.y; r[i + 2] += bv.z; r[i + 3] += bv.w;
          }
        }
        if (epi & RTTS_EPI_RELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = fmaxf(r[i], 0.f);
        }
        if (epi & RTTS_EPI_GATE) {
          // gate rows [32 x 64 B] -> staging (coalesced: 4 lanes per row, 8 rows per instruction) -> this thread's row.  The tile is
          // shared with the output rows: with a bf16 output the second chunk of a pair must not touch the half of each row that
          // already holds the first chunk's results, so the gate goes where that chunk's own results will go.
          const int ghalf = (epi & RTTS_EPI_OUT_BF16) ? ((c0 >> 5) & 1) * 4 : 0;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int gr = it * 8 + (lane >> 2), gc = lane & 3;
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(p.gate + static_cast<int64_t>(m0 + quarter * 32 + gr) * p.ldgate + col) + gc);
            sts128(gstage + gr * 128 + (((ghalf + gc) ^ (gr & 7)) << 4), u);
          }
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 u = lds128(gstage + lane * 128 + (((ghalf + q) ^ (lane & 7)) << 4));
            const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (!(bf16_lo(wv[e]) > 0.f)) r[q * 8 + 2 * e] = 0.f;
              if (!(bf16_hi(wv[e]) > 0.f)) r[q * 8 + 2 * e + 1] = 0.f;
            }
          }
        }
        if (epi & RTTS_EPI_GATE) __syncwarp();       // every lane has read its gate row: the staging tile is free again
        if (epi & RTTS_EPI_OUT_BF16) {
          // two 32-column chunks make one 128-byte row segment
          const int half = (c0 >> 5) & 1;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 u;
            u.x = pack_bf16(r[q * 8 + 0], r[q * 8 + 1]); u.y = pack_bf16(r[q * 8 + 2], r[q * 8 + 3]);
            u.z = pack_bf16(r[q * 8 + 4], r[q * 8 + 5]); u.w = pack_bf16(r[q * 8 + 6], r[q * 8 + 7]);
            sts128(stage + lane * 128 + (((half * 4 + q) ^ (lane & 7)) << 4), u);
          }
          if (half == 1) {
            __syncwarp();
            __nv_bfloat16* base = static_cast<__nv_bfloat16*>(p.C) + static_cast<int64_t>(m0 + quarter * 32) * p.ldc + (col - 32);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rr = it * 4 + srow;
              reinterpret_cast<uint4*>(base + static_cast<int64_t>(rr) * p.ldc)[sch] = lds128(stage + rr * 128 + ((sch ^ (rr & 7)) << 4));
            }
            __syncwarp();
          }
        } else if (epi & RTTS_EPI_ATOMIC) {
          float* dst = static_cast<float*>(p.C) + static_cast<int64_t>(row) * p.ldc + col;
#pragma unroll
          for (int i = 0; i < 32; i += 4)      // 16-byte vector reduction (sm_90+): a quarter of the atomic instructions
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(r[i]), "f"(r[i + 1]), "f"(r[i + 2]), "f"(r[i + 3]) : "memory");
        } else {
          // 32 fp32 columns = one 128-byte row segment
#pragma unroll
          for (int q = 0; q < 8; ++q)
            sts128(stage + lane * 128 + ((q ^ (lane & 7)) << 4),
                   make_uint4(__float_as_uint(r[q * 4]), __float_as_uint(r[q * 4 + 1]), __float_as_uint(r[q * 4 + 2]), __float_as_uint(r[q * 4 + 3])));
          // residual rows / keep-mask bytes of this chunk in the layout of the store phase (coalesced); they were pulled into L2 at the
          // start of the tile, and r[] is dead here, so the eight loads in flight cost no extra registers
          float4 rv[8];
          uint32_t kw[8];
          if (epi & (RTTS_EPI_RESID_ADD | RTTS_EPI_RESID_SUB)) {
            const float* rbase = reinterpret_cast<const float*>(p.gate) + static_cast<int64_t>(m0 + quarter * 32) * p.ldgate + col;
#pragma unroll
            for (int it = 0; it < 8; ++it) rv[it] = *reinterpret_cast<const float4*>(rbase + static_cast<int64_t>(it * 4 + srow) * p.ldgate + sch * 4);
          }
          if (p.keep != nullptr) {
            const uint8_t* kbase = p.keep + static_cast<int64_t>(m0 + quarter * 32) * p.ldkeep + col;
#pragma unroll
            for (int it = 0; it < 8; ++it) kw[it] = *reinterpret_cast<const uint32_t*>(kbase + static_cast<int64_t>(it * 4 + srow) * p.ldkeep + sch * 4);
          }
          __syncwarp();
          float* base = static_cast<float*>(p.C) + static_cast<int64_t>(m0 + quarter * 32) * p.ldc + col;
          if ((epi & (RTTS_EPI_RESID_ADD | RTTS_EPI_RESID_SUB)) || p.keep != nullptr) {
            // dropout and residual stream fused into the store: C = [resid +/-] keep * scale * result (reversible forward: +, reconstruction
            // of the block input in the reversible backward: -).  Same coalesced 16-byte accesses as the store; C may alias resid.
            const bool has_resid = (epi & (RTTS_EPI_RESID_ADD | RTTS_EPI_RESID_SUB)) != 0;
            const float sgn = (epi & RTTS_EPI_RESID_SUB) ? -1.f : 1.f;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rr = it * 4 + srow;
              const uint4 u = lds128(stage + rr * 128 + ((sch ^ (rr & 7)) << 4));
              float4 o = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
              if (p.keep != nullptr) {
                o.x = (kw[it] & 0x000000ffu) ? o.x * p.keep_scale : 0.f; o.y = (kw[it] & 0x0000ff00u) ? o.y * p.keep_scale : 0.f;
                o.z = (kw[it] & 0x00ff0000u) ? o.z * p.keep_scale : 0.f; o.w = (kw[it] & 0xff000000u) ? o.w * p.keep_scale : 0.f;
              }
              if (has_resid) {
                o.x = fmaf(sgn, o.x, rv[it].x); o.y = fmaf(sgn, o.y, rv[it].y);
                o.z = fmaf(sgn, o.z, rv[it].z); o.w = fmaf(sgn, o.w, rv[it].w);
              }
              *reinterpret_cast<float4*>(base + static_cast<int64_t>(rr) * p.ldc + sch * 4) = o;
            }
          } else {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rr = it * 4 + srow;
              reinterpret_cast<uint4*>(base + static_cast<int64_t>(rr) * p.ldc)[sch] = lds128(stage + rr * 128 + ((sch ^ (rr & 7)) << 4));
            }
          }
          __syncwarp();
        }
        if (epi & RTTS_EPI_COLSUM) {
          const float tot = warp_column_sums(r, lane);
          atomicAdd(p.colsum + col + lane, tot);
        }
      }
      tc_fence_before_sync();
      mbar_arrive(acc_empty + acc);            // 128 arrivals: the MMA warp may overwrite this accumulator
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------
static int make_tmap(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, int64_t ld, uint32_t box_inner,
                     uint32_t box_outer) {
  return make_tmap_bf16(map, base, inner, outer, ld, box_inner, box_outer);
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
    if (e != cudaSuccess) return fail(kErrCuda, "rtts_gemm_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int grid = p.work < kNumSMs ? p.work : kNumSMs;
  gemm_bf16_kernel<BN, A_MN, B_MN><<<grid, gemm_threads(BN), Cfg::kSmem, stream>>>(ta, tb, p);
  return check_launch("rtts_gemm_bf16");
}

template <int BN>
static int dispatch_gemm(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, GemmParams& p, int K_total, cudaStream_t s) {
  CUtensorMap ta, tb;
  int rc;
  if (a_mn) rc = make_tmap(&ta, A, p.M, K_total, lda, 64, kBK);
  else rc = make_tmap(&ta, A, K_total, p.M, lda, kBK, kBM);
  if (rc) return rc;
  if (b_mn) rc = make_tmap(&tb, B, p.N, K_total, ldb, 64, kBK);
  else rc = make_tmap(&tb, B, K_total, p.N, ldb, kBK, 128);
  if (rc) return rc;
  p.tiles_n = p.N / BN;
  p.tiles_mn = (p.M / kBM) * p.tiles_n;
  if (a_mn && b_mn) return launch_gemm<BN, true, true>(ta, tb, p, s);
  if (a_mn) return launch_gemm<BN, true, false>(ta, tb, p, s);
  if (b_mn) return launch_gemm<BN, false, true>(ta, tb, p, s);
  return launch_gemm<BN, false, false>(ta, tb, p, s);
}

}  // namespace rtts

using namespace rtts;

extern "C" int rtts_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major, void* C,
                              int64_t ldc, const float* bias, const void* gate, int64_t ldgate, float* colsum, int M, int N,
                              int K, int epilogue, int split_k, void* stream) {
  return rtts_gemm_bf16_dropout(A, lda, a_mn_major, B, ldb, b_mn_major, C, ldc, bias, gate, ldgate, colsum, nullptr, 0, 1.f, M, N, K, epilogue,
                                split_k, stream);
}

extern "C" int rtts_gemm_bf16_dropout(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major, void* C,
                                      int64_t ldc, const float* bias, const void* gate, int64_t ldgate, float* colsum,
                                      const uint8_t* keep_mask, int64_t ldkeep, float keep_scale, int M, int N, int K, int epilogue,
                                      int split_k, void* stream) {
  RTTS_REQUIRE(!keep_mask || (!(epilogue & (RTTS_EPI_OUT_BF16 | RTTS_EPI_ATOMIC | RTTS_EPI_COLSUM)) && ldkeep % 4 == 0 &&
                              (reinterpret_cast<uintptr_t>(keep_mask) & 3) == 0),
               "rtts_gemm_bf16_dropout: the keep mask needs an fp32 non-atomic output without column sums, 4-byte aligned rows");
  RTTS_REQUIRE(A && B && C, "rtts_gemm_bf16: null pointer");
  RTTS_REQUIRE(M > 0 && N > 0 && K > 0 && M % kBM == 0 && N % 128 == 0, "rtts_gemm_bf16: M=%d, N=%d must be multiples of 128", M, N);
  RTTS_REQUIRE(split_k >= 1 && K % (kBK * split_k) == 0, "rtts_gemm_bf16: K=%d must be a multiple of 64*split_k", K);
  RTTS_REQUIRE(split_k == 1 || (epilogue & RTTS_EPI_ATOMIC), "rtts_gemm_bf16: split_k > 1 needs RTTS_EPI_ATOMIC");
  RTTS_REQUIRE(!((epilogue & RTTS_EPI_ATOMIC) && (epilogue & (RTTS_EPI_OUT_BF16 | RTTS_EPI_BIAS | RTTS_EPI_RELU | RTTS_EPI_GATE | RTTS_EPI_COLSUM))),
               "rtts_gemm_bf16: RTTS_EPI_ATOMIC cannot be combined with other epilogue flags");
  RTTS_REQUIRE(!(epilogue & RTTS_EPI_BIAS) || bias, "rtts_gemm_bf16: bias flag without bias pointer");
  RTTS_REQUIRE(!(epilogue & RTTS_EPI_GATE) || gate, "rtts_gemm_bf16: gate flag without gate pointer");
  constexpr int kResid = RTTS_EPI_RESID_ADD | RTTS_EPI_RESID_SUB;
  RTTS_REQUIRE(!(epilogue & kResid) || (gate && !(epilogue & (RTTS_EPI_GATE | RTTS_EPI_OUT_BF16 | RTTS_EPI_ATOMIC)) && (epilogue & kResid) != kResid &&
                                        ldgate % 4 == 0 && (reinterpret_cast<uintptr_t>(gate) & 15) == 0),
               "rtts_gemm_bf16: a residual needs its fp32 pointer in `gate` (16-byte aligned), an fp32 non-atomic output and one of ADD / SUB");
  RTTS_REQUIRE(!(epilogue & RTTS_EPI_COLSUM) || colsum, "rtts_gemm_bf16: colsum flag without colsum pointer");
  RTTS_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0 && (ldgate % 8 == 0), "rtts_gemm_bf16: leading dimensions must be multiples of 8");
  RTTS_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15) == 0,
               "rtts_gemm_bf16: operands must be 16-byte aligned");
  GemmParams p;
  p.C = C; p.ldc = ldc; p.bias = bias; p.gate = static_cast<const __nv_bfloat16*>(gate); p.ldgate = ldgate; p.colsum = colsum;
  p.keep = keep_mask; p.ldkeep = ldkeep; p.keep_scale = keep_scale;
  p.M = M; p.N = N; p.K = K / split_k; p.epilogue = epilogue;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // 128x256 tiles halve the shared-memory traffic per flop; use them when they still give every SM several tiles
  const int64_t work256 = N % 256 == 0 ? static_cast<int64_t>(M / kBM) * (N / 256) * split_k : 0;
  if (work256 >= 4 * kNumSMs) {
    p.work = static_cast<int>(work256);
    return dispatch_gemm<256>(A, lda, a_mn_major, B, ldb, b_mn_major, p, K, s);
  }
  p.work = (M / kBM) * (N / 128) * split_k;
  return dispatch_gemm<128>(A, lda, a_mn_major, B, ldb, b_mn_major, p, K, s);
}
