// LSH hashing on the tensor pipe (rtts_lsh_hash_tc): the random-rotation projections of rp R2 / hf:717-758 as tcgen05 MMAs that
// reproduce the fp32 products EXACTLY, followed by the argmax over cat([r, -r]) in the epilogue.
//
// Why this is exact enough for bit-equal bucket ids.  qk is bf16 (8 significant bits).  Each fp32 rotation entry w is split into
// three bf16 numbers hi + mid + lo == w (24 = 3 x 8 significant bits; the residuals w - hi and w - hi - mid are exact in fp32), so
// every product x * part has at most 16 significant bits and is formed exactly by the bf16 MMA; only the fp32 accumulation of the
// 3 x 64 products rounds, as the fp32 FMA chain of the reference does (in another order).  The parts are accumulated smallest
// first.  The projections are therefore fp32-accurate dot products of the SAME operands the reference multiplies, and the bucket
// id differs from an exact (fp64) evaluation only where the top-2 margin is within fp32 rounding of the row norm - the same
// caveat the reference's own fp32 einsum carries (tests/test_kernels_gpu.py::test_hash_bit_exact states the margin).
//
// Work decomposition.  A tile = 128 consecutive tokens of one (batch, head): A = the [128 x 64] bf16 rows (K-major, one TMA box out
// of the token-major qk tensor), B = the split rotations [P x 64] x 3 (K-major, resident in shared memory for the whole CTA:
// they depend on the head only in the per-head variant, where blockIdx.y selects the head), D = [128 x P] fp32 in TMEM, P = R *
// n_buckets / 2 <= 256 projections.  Persistent CTAs, warp roles: TMA loader / MMA issuer / 16 epilogue warps (thread = token x share
// of the hash rounds: reads its projections from TMEM, running (max, first index) / (min, first index) per hash round with torch.argmax's first-maximum
// rule, writes R bucket ids; also |x|^2 of its row for the attention kernels' key normalisation).  Two TMEM accumulators, three
// A stages.  Algorithmic traffic: B*T*D*2 bytes read + B*H*R*T*4 bytes written (HBM-bound), 2*3*B*T*D*P flop on the tensor pipe.
#include "common.cuh"
#include "host_util.h"
#include "rtts_b200.h"
#include "tma_host.h"

namespace rtts {

constexpr int kHtEpiWarps = 16;        // four warps per TMEM lane quarter: each takes a contiguous share of the hash rounds
constexpr int kHtThreads = 64 + kHtEpiWarps * 32;      // warp 0: TMA, warp 1: MMA, warps 2-17: epilogue
constexpr int kHtStages = 3;
constexpr int kHtTileBytes = 128 * 128;
constexpr int kHtMaxP = 256;

struct HashTcSmem {
  static constexpr int kOffA = 0;                                   // kHtStages x 16 KB
  static constexpr int kOffB = kHtStages * kHtTileBytes;            // 3 parts x (P <= 256 rows x 128 B)
  static constexpr int kOffBar = kOffB + 3 * kHtMaxP * 128;         // full[3], empty[3], acc_full[2], acc_free[2]
  static constexpr int kOffTmem = kOffBar + 10 * 8;
  static constexpr int kTotal = kOffTmem + 8;
};

// rot fp32 [rot_heads][64][P]  ->  ws bf16 [rot_heads][3 parts: lo, mid, hi][P][64]
// (P projections starting at column p0 of the P_all = R * n_buckets / 2 the rotation tensor holds: one group of hash rounds)
__global__ void __launch_bounds__(256) lsh_hash_split_kernel(const float* __restrict__ rot, __nv_bfloat16* __restrict__ ws, int P, int P_all, int p0,
                                                             int total) {
  const int idx = blockIdx.x * 256 + threadIdx.x;      // (hr, p, k), k fastest
  if (idx >= total) return;
  const int k = idx & 63, p = (idx >> 6) % P, hr = (idx >> 6) / P;
  const float w = rot[(static_cast<int64_t>(hr) * 64 + k) * P_all + p0 + p];
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const float e1 = w - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(e1);
  const float e2 = e1 - __bfloat162float(mid);
  const __nv_bfloat16 lo = __float2bfloat16_rn(e2);
  __nv_bfloat16* dst = ws + (static_cast<int64_t>(hr) * 3 * P + p) * 64 + k;
  dst[0] = lo;
  dst[static_cast<int64_t>(P) * 64] = mid;
  dst[static_cast<int64_t>(2) * P * 64] = hi;
}

struct HashTcParams {
  const __nv_bfloat16* ws;
  const uint8_t* pad_mask;
  int32_t* buckets;
  float* sumsq;
  int T, H, R, n_buckets, use_pad_bucket, rot_heads;      // R = rounds of THIS launch
  int r_base, R_all;      // first round of this launch, rounds of the bucket tensor
  int tiles_per_seq;      // T / 128
  int num_tiles;          // tiles this grid row (blockIdx.y) walks: B*H*T/128 (shared rotations) or B*T/128 (per head)
};

__global__ void __launch_bounds__(kHtThreads, 1) lsh_hash_mma_kernel(const __grid_constant__ CUtensorMap tmap_qk, const HashTcParams p) {
  using L = HashTcSmem;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* full = bars;            // [3] TMA -> MMA, epilogue
  uint64_t* empty = bars + 3;       // [3] tcgen05.commit + every epilogue warp
  uint64_t* acc_full = bars + 6;    // [2] tcgen05.commit
  uint64_t* acc_free = bars + 8;    // [2] every epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = p.n_buckets >> 1, P = p.R * half;
  const int hr = p.rot_heads == 1 ? 0 : blockIdx.y;

  if (tid == 0) {
    for (int s = 0; s < kHtStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1 + kHtEpiWarps);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full + a, 1);
      mbar_init(acc_free + a, kHtEpiWarps);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmap_qk);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  // the split rotations of this CTA's head: three K-major SWIZZLE_128B tiles of P rows
  {
    const __nv_bfloat16* src = p.ws + static_cast<int64_t>(hr) * 3 * P * 64;
    const uint32_t sB = smem_u32(smem + L::kOffB);
    for (int i = tid; i < 3 * P * 8; i += kHtThreads) {
      const int part = i / (P * 8), rem = i - part * (P * 8), row = rem >> 3, c = rem & 7;
      cp_async16(sB + part * (kHtMaxP * 128) + sw128_offset(row, c), src + (static_cast<int64_t>(part) * P + row) * 64 + c * 8);
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_proxy_async_smem();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (*tmem_slot != 0) __trap();
  constexpr uint32_t tmem = 0;

  // tile i of this CTA = global tile blockIdx.x + i * gridDim.x; (row of the [B*T, H*64] view, head) of a tile
  auto tile_coords = [&](int g, int& row0, int& h) {
    if (p.rot_heads == 1) {       // g = ((b * H) + h) * tiles_per_seq + tb
      const int tb = g % p.tiles_per_seq, bh = g / p.tiles_per_seq;
      h = bh % p.H;
      row0 = (bh / p.H) * p.T + tb * 128;
    } else {                      // g = b * tiles_per_seq + tb, head = blockIdx.y
      h = blockIdx.y;
      row0 = (g / p.tiles_per_seq) * p.T + (g % p.tiles_per_seq) * 128;
    }
  };

  if (warp == 0) {
    // ================================================= TMA loader =================================================
    if (elect_one()) {
      int i = 0;
      for (int g = blockIdx.x; g < p.num_tiles; g += gridDim.x, ++i) {
        const int s = i % kHtStages;
        mbar_wait(empty + s, ((i / kHtStages) & 1) ^ 1);
        int row0, h;
        tile_coords(g, row0, h);
        mbar_arrive_expect_tx(full + s, kHtTileBytes);
        tma_load_2d(smem_u32(smem + L::kOffA + s * kHtTileBytes), &tmap_qk, full + s, h * 64, row0);
      }
    }
  } else if (warp == 1) {
    // ================================================= MMA issuer =================================================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, P, false, false);
      constexpr uint32_t hi = umma_desc_hi_sw128(1024);
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(smem + L::kOffB), 16);
      int i = 0;
      for (int g = blockIdx.x; g < p.num_tiles; g += gridDim.x, ++i) {
        const int s = i % kHtStages, a = i & 1;
        mbar_wait(full + s, (i / kHtStages) & 1);
        mbar_wait(acc_free + a, ((i >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t a_lo = umma_desc_lo(smem_u32(smem + L::kOffA + s * kHtTileBytes), 16);
#pragma unroll
        for (int part = 0; part < 3; ++part)        // lo, mid, hi: the smallest terms enter the fp32 accumulator first
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss_lo(tmem + a * 256, a_lo + kk * 2, b_lo0 + part * ((kHtMaxP * 128) >> 4) + kk * 2, hi, idesc, (part | kk) != 0);
        umma_commit(acc_full + a);
        umma_commit(empty + s);
      }
    }
  } else {
    // ================================================= epilogue ===================================================
    const int q = warp & 3, row = q * 32 + lane;       // TMEM lane quarter of this warp, token row inside the tile
    // column groups: G warps per lane quarter split the R rounds (whole rounds, chunks of 16 columns); the others only keep the barriers going
    int G = 4;
    while (G > 1 && (p.R % G != 0 || (P / G) % 16 != 0)) G >>= 1;
    const int cg = (warp - 2) >> 2;
    const bool active = cg < G;
    const int col_lo = active ? cg * (P / G) : 0, col_hi = active ? col_lo + P / G : 0, r_first = active ? cg * (p.R / G) : 0;
    const uint32_t t_row = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const int stride = p.use_pad_bucket ? p.n_buckets + 1 : p.n_buckets;
    int i = 0;
    for (int g = blockIdx.x; g < p.num_tiles; g += gridDim.x, ++i) {
      const int s = i % kHtStages, a = i & 1;
      int row0, h;
      tile_coords(g, row0, h);
      const int b = row0 / p.T, t = row0 - b * p.T + row;
      const int64_t bh = static_cast<int64_t>(b) * p.H + h;
      mbar_wait(full + s, (i / kHtStages) & 1);         // the rows are in shared memory (|x|^2 is taken from there)
      if (p.sumsq != nullptr && cg == 0) {
        const uint32_t sA = smem_u32(smem + L::kOffA + s * kHtTileBytes);
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {      // same element order as the fp32 kernel: k = 0 .. 63
          const uint4 u = lds128(sA + sw128_offset(row, c));
          ss = fmaf(bf16_lo(u.x), bf16_lo(u.x), ss); ss = fmaf(bf16_hi(u.x), bf16_hi(u.x), ss);
          ss = fmaf(bf16_lo(u.y), bf16_lo(u.y), ss); ss = fmaf(bf16_hi(u.y), bf16_hi(u.y), ss);
          ss = fmaf(bf16_lo(u.z), bf16_lo(u.z), ss); ss = fmaf(bf16_hi(u.z), bf16_hi(u.z), ss);
          ss = fmaf(bf16_lo(u.w), bf16_lo(u.w), ss); ss = fmaf(bf16_hi(u.w), bf16_hi(u.w), ss);
        }
        p.sumsq[bh * p.T + t] = ss;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);            // this warp is done with the stage
      const bool padded = p.use_pad_bucket && p.pad_mask != nullptr && p.pad_mask[static_cast<int64_t>(b) * p.T + t] == 0;
      int32_t* out = p.buckets + (bh * p.R_all + p.r_base) * p.T + t;
      mbar_wait(acc_full + a, (i >> 1) & 1);
      tc_fence_after_sync();
      float vmax = 0.f, vmin = 0.f;
      int imax = 0, imin = 0, idx = 0, r = r_first;    // idx = projection index inside the round
#pragma unroll 1
      for (int c0 = col_lo; c0 < col_hi; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(t_row + a * 256 + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float acc = __uint_as_float(v[j]);
          if (idx == 0) { vmax = acc; vmin = acc; imax = 0; imin = 0; }
          else {
            if (acc > vmax) { vmax = acc; imax = idx; }
            if (acc < vmin) { vmin = acc; imin = idx; }
          }
          if (++idx == half) {
            // torch.argmax returns the FIRST maximum of cat([r, -r]): every +r precedes every -r, so a tie goes to +r
            int id = (vmax >= -vmin) ? imax : half + imin;
            if (padded) id = p.n_buckets;
            out[static_cast<int64_t>(r) * p.T] = (p.r_base + r) * stride + id;
            idx = 0;
            ++r;
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free + a);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace rtts

using namespace rtts;

// Rounds per launch: all of them when their projections fit one accumulator (<= 256 columns), else the largest divisor of R that does.
static int rounds_per_launch(int R, int n_buckets) {
  const int half = n_buckets / 2;
  if (half < 1 || half > kHtMaxP) return 0;
  int g = R;
  while (g > 1 && (g * half > kHtMaxP || R % g != 0)) --g;
  return g * half <= kHtMaxP ? g : 0;
}

extern "C" int64_t rtts_lsh_hash_tc_workspace_bytes(int rot_heads, int R, int n_buckets) {
  const int g = rounds_per_launch(R, n_buckets);
  return static_cast<int64_t>(rot_heads) * 3 * (g > 0 ? g : R) * (n_buckets / 2) * 64 * 2;
}

extern "C" int rtts_lsh_hash_tc_supported(int T, int dh, int R, int n_buckets) {
  const int g = rounds_per_launch(R, n_buckets);
  const int P = g * (n_buckets / 2);      // projections per launch
  return dh == 64 && T % 128 == 0 && n_buckets % 2 == 0 && g > 0 && P % 16 == 0 && P >= 16;
}

extern "C" int rtts_lsh_hash_tc(const void* qk, int64_t ld, const float* rot, int rot_heads, const uint8_t* pad_mask, int use_pad_bucket,
                                int32_t* buckets, float* sumsq, void* workspace, int B, int T, int H, int dh, int R, int n_buckets,
                                void* stream) {
  RTTS_REQUIRE(qk && rot && buckets && workspace, "rtts_lsh_hash_tc: null pointer");
  RTTS_REQUIRE(rtts_lsh_hash_tc_supported(T, dh, R, n_buckets), "rtts_lsh_hash_tc: unsupported shape (dh 64, T %% 128, projections per launch a multiple of 16)");
  RTTS_REQUIRE(rot_heads == 1 || rot_heads == H, "rtts_lsh_hash_tc: rot_heads must be 1 or H");
  RTTS_REQUIRE(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(qk) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
               "rtts_lsh_hash_tc: qk / workspace must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUtensorMap tmap;
  int rc = make_tmap_bf16(&tmap, qk, static_cast<uint64_t>(H) * 64, static_cast<uint64_t>(B) * T, ld, 64, 128);
  if (rc != kOk) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(lsh_hash_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HashTcSmem::kTotal);
    if (e != cudaSuccess) return fail(kErrCuda, "rtts_lsh_hash_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int half = n_buckets / 2, P_all = R * half;
  const int g = rounds_per_launch(R, n_buckets), P = g * half;
  HashTcParams p;
  p.ws = static_cast<const __nv_bfloat16*>(workspace);
  p.pad_mask = pad_mask; p.buckets = buckets;
  p.T = T; p.H = H; p.R = g; p.R_all = R; p.n_buckets = n_buckets; p.use_pad_bucket = use_pad_bucket; p.rot_heads = rot_heads;
  p.tiles_per_seq = T / 128;
  p.num_tiles = rot_heads == 1 ? B * H * p.tiles_per_seq : B * p.tiles_per_seq;
  const int per_row = rot_heads == 1 ? kNumSMs : (kNumSMs / H > 0 ? kNumSMs / H : 1);
  dim3 grid(p.num_tiles < per_row ? p.num_tiles : per_row, rot_heads == 1 ? 1 : H);
  // one launch per group of g rounds (all R at once unless R * n_buckets / 2 > 256); the workspace is reused in stream order
  for (int r0 = 0; r0 < R; r0 += g) {
    const int total = rot_heads * P * 64;
    lsh_hash_split_kernel<<<(total + 255) / 256, 256, 0, s>>>(rot, static_cast<__nv_bfloat16*>(workspace), P, P_all, r0 * half, total);
    rc = check_launch("rtts_lsh_hash_tc (split)");
    if (rc != kOk) return rc;
    p.r_base = r0;
    p.sumsq = r0 == 0 ? sumsq : nullptr;
    lsh_hash_mma_kernel<<<grid, kHtThreads, HashTcSmem::kTotal, s>>>(tmap, p);
    rc = check_launch("rtts_lsh_hash_tc");
    if (rc != kOk) return rc;
  }
  return kOk;
}
