// Dense decoder -> encoder attention core, forward and backward (rtts_xattn_fwd / rtts_xattn_bwd): softmax(Q K^T / sqrt(dh)) V
// over the <= 256 encoder positions of nn.MultiheadAttention (ref:reformer_tts/model/reformer.py:161-186), with the key padding
// mask and the attention-probability dropout of the module inside the kernel.
//
// The problem is tiny per (batch, head) - 1024 x 256 scores - so the kernels are NOT persistent pipelines: one CTA per
// (batch, head, 128-query tile), everything of the tile resident (Q tile, all K / V rows of the head in SWIZZLE_128B shared-memory
// tiles, the whole score row of a query in TMEM), and the SM overlaps CTAs instead of stages.
//   forward   S = Q K^T (tcgen05, M128 N=S K64) -> TMEM;  thread = query row: max, exp2, row sum, dropout, P (bf16 pairs) written back
//             over S;  O = P V (A from TMEM);  out = O / (rowsum (1-p)) -> bf16 rows of the token-major [B, T, D] output;  lse saved.
//   backward  S and dP = dO V^T recomputed into TMEM (512 columns);  thread = (query row, half of the keys): P = exp2(s - lse),
//             dS = P o (D o dP - delta) scale with D the regenerated dropout mask;  dS and D o P (bf16) to shared memory, where ONE tile
//             serves as the K-major A operand of dQ = dS K and as the MN-major A operand of dK = dS^T Q (dV = (D o P)^T dO alike);
//             dQ rows stored as bf16, dK / dV partial sums of the tile added to the fp32 [B, S, 2D] gradient with vector reductions.
// Dropout: a counter-based hash of (seed, element index) - the seed is a device word drawn from torch's generator at the point of the
// call order where nn.MultiheadAttention would draw, so Deterministic's RNG replay (and a CUDA-graph replay) regenerate the same mask
// in the forward, the reversible recompute and the backward without anything being stored.
#include <cfloat>

#include "common.cuh"
#include "host_util.h"
#include "rtts_b200.h"

namespace rtts {
namespace xa {

constexpr int kDh = 64;
constexpr int kQ = 128;                 // queries per CTA
constexpr int kMaxS = 256;              // keys (all of them resident)
constexpr float kLog2e = 1.4426950408889634f;

struct Params {
  long long* trace;                             // debug: clock64 stamps of CTA (0,0,0) (nullable)
  const __nv_bfloat16* q;  int64_t ldq;          // [B, T, H*64]
  const __nv_bfloat16* k;  const __nv_bfloat16* v;  int64_t ldkv;      // [B, S, ...]: head h at column h*64 of each
  const uint8_t* keep;                           // [B, S] 1 = valid key (nullable)
  const __nv_bfloat16* dout;  int64_t lddo;      // backward: [B, T, H*64]
  __nv_bfloat16* out;  int64_t ldo;              // forward: [B, T, H*64]
  float* lse;                                    // [B, H, T]  log2-domain: max * c + log2(sum)
  const float* delta;                            // backward: [B, H, T]  <dout, out>
  __nv_bfloat16* dq;  int64_t lddq;              // backward: [B, T, H*64]
  float* dk;  float* dv;  int64_t lddkv;         // backward: fp32 [B, S, ...] accumulated (head h at column h*64)
  const uint64_t* seed;                          // device word (nullable: no dropout)
  uint32_t drop_thr;                             // drop when the 16-bit hash half < drop_thr  (p * 65536)
  float keep_scale;                              // 1 / (1 - p)
  float c;                                       // softmax scale * log2(e)
  float scale;                                   // softmax scale
  int B, T, S, H;
};

// Two dropout decisions per hash: element pair (s, s+1), s even, of query row `row` (global row index (b*H + h)*T + t).
__device__ __forceinline__ uint32_t drop_hash(uint32_t seed_lo, uint32_t seed_hi, uint32_t row, uint32_t pair, uint32_t half_s) {
  uint32_t x = (row * half_s + pair) * 0x9E3779B1u + seed_lo;
  x ^= x >> 15;
  x *= 0x85EBCA77u;
  x ^= x >> 13;
  x = x * 0xC2B2AE3Du + seed_hi;
  x ^= x >> 16;
  return x;
}

#ifdef RTTS_TRACE
#define XA_STAMP(k) do { if (p.trace != nullptr && blockIdx.x + blockIdx.y + blockIdx.z == 0 && threadIdx.x == 0) p.trace[k] = clock64(); } while (0)
#else
#define XA_STAMP(k) do { } while (0)
#endif

// rows x 128 B (row pitch ld elements) -> SWIZZLE_128B tile, 16-byte cp.async pieces.  Thread t owns piece t % 8 of the rows t / 8,
// t / 8 + nthreads / 8, ...: nthreads / 8 is a multiple of 8, so the swizzle term of its piece never changes and both addresses advance
// by a constant per copy (the index form - shifts, masks, a 64-bit multiply per piece - took 2 100 of a forward CTA's 14 200 cycles).
__device__ __forceinline__ void load_rows(uint32_t s_tile, const __nv_bfloat16* g, int64_t ld, int rows, int tid, int nthreads) {
  const int r0 = tid >> 3, c = tid & 7, step = nthreads >> 3;
  uint32_t dst = s_tile + sw128_offset(r0, c);
  const char* src = reinterpret_cast<const char*>(g + static_cast<int64_t>(r0) * ld + c * 8);
  const int64_t src_step = static_cast<int64_t>(step) * ld * 2;
  const uint32_t dst_step = static_cast<uint32_t>(step) * 128u;
#pragma unroll 4
  for (int r = r0; r < rows; r += step) {
    cp_async16(dst, src);
    dst += dst_step;
    src += src_step;
  }
}

__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 18); ++spin) {      // 2^18 x 20 us > 5 s: a protocol bug traps instead of hanging the GPU
    if (mbar_try_wait_suspend(bar, parity, 20000u)) return;
  }
  __trap();
}

// ------------------------------------------------------------------------------------------------------------ forward
struct FwdSmem {
  static constexpr int kOffQ = 0;                           // 16 KB
  static constexpr int kOffK = kQ * 128;                    // 32 KB
  static constexpr int kOffV = kOffK + kMaxS * 128;         // 32 KB
  static constexpr int kOffBias = kOffV + kMaxS * 128;      // float[256]: 0 / -inf per key
  static constexpr int kOffXch = kOffBias + kMaxS * 4;      // float[2][128], then uint8[16]: chunk of 16 keys holds a padded key
  static constexpr int kOffBar = kOffXch + 2 * kQ * 4 + 16;
  static constexpr int kTotal = kOffBar + 32;
};

constexpr int kFwdThreads = 256;        // thread = (query row, half of the keys, half of the output row)
__global__ void __launch_bounds__(kFwdThreads, 2) xattn_fwd_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + FwdSmem::kOffBar + 16);
  float* bias = reinterpret_cast<float*>(smem + FwdSmem::kOffBias);
  float* xch = reinterpret_cast<float*>(smem + FwdSmem::kOffXch);      // [2 halves][128 rows]: row maxima, then row sums
  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x, b = blockIdx.y, t0 = blockIdx.z * kQ;
  const int S = p.S;

  XA_STAMP(0);
  if (tid == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  load_rows(sbase + FwdSmem::kOffQ, p.q + (static_cast<int64_t>(b) * p.T + t0) * p.ldq + h * kDh, p.ldq, kQ, tid, kFwdThreads);
  load_rows(sbase + FwdSmem::kOffK, p.k + static_cast<int64_t>(b) * S * p.ldkv + h * kDh, p.ldkv, S, tid, kFwdThreads);
  load_rows(sbase + FwdSmem::kOffV, p.v + static_cast<int64_t>(b) * S * p.ldkv + h * kDh, p.ldkv, S, tid, kFwdThreads);
  cp_async_commit();
  XA_STAMP(1);
  for (int j = tid; j < S; j += kFwdThreads) bias[j] = (p.keep == nullptr || p.keep[static_cast<int64_t>(b) * S + j]) ? 0.f : -INFINITY;
  cp_async_wait<0>();
  __syncthreads();
  // chunks of 16 keys without a padded key skip the bias (one shared-memory load + add per score otherwise)
  if (tid < S / 16) {
    float acc = 0.f;
    for (int i = 0; i < 16; ++i) acc += bias[tid * 16 + i];
    reinterpret_cast<uint8_t*>(smem + FwdSmem::kOffXch)[2 * kQ * 4 + tid] = acc < 0.f;
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  XA_STAMP(2);
  constexpr uint32_t hi = umma_desc_hi_sw128(1024);

  if (warp == 0 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(kQ, S, false, false);
    const uint32_t a_lo = umma_desc_lo(sbase + FwdSmem::kOffQ, 16), b_lo = umma_desc_lo(sbase + FwdSmem::kOffK, 16);
#pragma unroll
    for (int kk = 0; kk < kDh / 16; ++kk) umma_ss_lo(tmem, a_lo + kk * 2, b_lo + kk * 2, hi, idesc, kk > 0);
    umma_commit(bars);
  }
  __syncwarp();
  bar_wait(bars, 0);
  tc_fence_after_sync();
  XA_STAMP(3);

  const int rowi = tid & (kQ - 1), half = tid >> 7;
  const uint8_t* padded = reinterpret_cast<const uint8_t*>(smem + FwdSmem::kOffXch) + 2 * kQ * 4;
  const int hs = S >> 1, k0 = half * hs;       // this thread's keys [k0, k0 + hs)
  const uint32_t t_row = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const uint32_t row = static_cast<uint32_t>((b * p.H + h) * p.T + t0 + rowi);
  float mx = -INFINITY;
  for (int c0 = k0; c0 < k0 + hs; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(t_row + c0, r);
    tmem_ld_wait();
    if (padded[c0 >> 4]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(r[i]) + bias[c0 + i]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
    }
  }
  XA_STAMP(4);
  xch[half * kQ + rowi] = mx;
  __syncthreads();
  mx = fmaxf(mx, xch[(half ^ 1) * kQ + rowi]);
  __syncthreads();                          // (the slots are reused for the sums)
  const float neg = -mx * p.c;
  uint32_t seed_lo = 0, seed_hi = 0;
  const bool drop = p.seed != nullptr;
  if (drop) {
    const uint64_t sd = *p.seed;
    seed_lo = static_cast<uint32_t>(sd);
    seed_hi = static_cast<uint32_t>(sd >> 32);
  }
  // P blocks are written back over the S columns of the SAME half that have been consumed: keys [k0, k0 + hs) -> columns k0 + (key - k0) / 2
  float sum = 0.f;
  for (int c0 = k0; c0 < k0 + hs; c0 += 16) {
    uint32_t r[16], pk[8];
    tmem_ld16(t_row + c0, r);
    tmem_ld_wait();
    if (padded[c0 >> 4]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + bias[c0 + i]);
    }
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      float e0 = exp2f(fmaf(__uint_as_float(r[i]), p.c, neg));
      float e1 = exp2f(fmaf(__uint_as_float(r[i + 1]), p.c, neg));
      sum += e0 + e1;                      // the normaliser is taken BEFORE the dropout (nn.MultiheadAttention drops normalised probabilities)
      if (drop) {
        const uint32_t hsh = drop_hash(seed_lo, seed_hi, row, static_cast<uint32_t>(c0 + i) >> 1, static_cast<uint32_t>(S) >> 1);
        if ((hsh & 0xffffu) < p.drop_thr) e0 = 0.f;
        if ((hsh >> 16) < p.drop_thr) e1 = 0.f;
      }
      pk[i >> 1] = pack_bf16(e0, e1);
    }
    tmem_st8(t_row + k0 + ((c0 - k0) >> 1), pk);
  }
  XA_STAMP(5);
  xch[half * kQ + rowi] = sum;
  tmem_st_wait();
  tc_fence_before_sync();
  __syncthreads();
  // O: 64 columns that hold neither half's P: behind P of half 0 when the halves are 128 keys wide, else behind all of S
  const uint32_t col_o = hs >= 128 ? static_cast<uint32_t>(hs >> 1) : static_cast<uint32_t>(S);
  if (warp == 0 && elect_one()) {
    tc_fence_after_sync();
    constexpr uint32_t idesc_o = umma_idesc_bf16(kQ, kDh, false, true);
    const uint32_t v_lo = umma_desc_lo(sbase + FwdSmem::kOffV, 0);
    for (int j = 0; j < S / 16; ++j) {
      const int key = 16 * j, hh = key >= hs;
      umma_ts_lo(tmem + col_o, tmem + hh * hs + ((key - hh * hs) >> 1), v_lo + j * (2048 >> 4), hi, idesc_o, j > 0);
    }
    umma_commit(bars + 1);
  }
  __syncwarp();
  sum += xch[(half ^ 1) * kQ + rowi];
  bar_wait(bars + 1, 0);
  tc_fence_after_sync();
  XA_STAMP(6);
  const float inv = p.keep_scale / sum;
  __nv_bfloat16* orow = p.out + (static_cast<int64_t>(b) * p.T + t0 + rowi) * p.ldo + h * kDh + half * 32;
  {
    uint32_t r[32];
    tmem_ld32(t_row + col_o + half * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16(__uint_as_float(r[8 * i + 0]) * inv, __uint_as_float(r[8 * i + 1]) * inv);
      u.y = pack_bf16(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv);
      u.z = pack_bf16(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv);
      u.w = pack_bf16(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv);
      reinterpret_cast<uint4*>(orow)[i] = u;
    }
  }
  if (half == 0) p.lse[row] = fmaf(mx, p.c, log2f(sum));
  tc_fence_before_sync();
  __syncthreads();
  XA_STAMP(7);
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------------------ backward
constexpr int kBwdThreads = 512;        // thread = (query row, quarter of the keys)
struct BwdSmem {
  static constexpr int kOffQ = 0;                           // 16 KB   Q tile      [128 queries x 64]
  static constexpr int kOffDO = kOffQ + kQ * 128;           // 16 KB   dO tile
  static constexpr int kOffK = kOffDO + kQ * 128;           // 32 KB   K           [S keys x 64]
  static constexpr int kOffV = kOffK + kMaxS * 128;         // 32 KB   V
  static constexpr int kAtom = kQ * 128;                    // 16 KB: 128 query rows x 64 keys (bf16), SWIZZLE_128B
  static constexpr int kOffDS = kOffV + kMaxS * 128;        // 64 KB   dS          4 atoms of 64 keys
  static constexpr int kOffDP = kOffDS + 4 * kAtom;         // 64 KB   D o P
  static constexpr int kOffBias = kOffDP + 4 * kAtom;       // float[256]
  static constexpr int kOffBar = kOffBias + kMaxS * 4;      // 2 barriers, the TMEM slot, uint8[16] chunk flags
  static constexpr int kTotal = kOffBar + 48;
  static_assert(kTotal <= 232448, "shared memory budget of one CTA (227 KB)");
};

__global__ void __launch_bounds__(kBwdThreads, 1) xattn_bwd_kernel(const Params p) {
  using L = BwdSmem;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffBar + 16);
  float* bias = reinterpret_cast<float*>(smem + L::kOffBias);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x, b = blockIdx.y, t0 = blockIdx.z * kQ;
  const int S = p.S;

  if (tid == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  load_rows(sbase + L::kOffQ, p.q + (static_cast<int64_t>(b) * p.T + t0) * p.ldq + h * kDh, p.ldq, kQ, tid, kBwdThreads);
  load_rows(sbase + L::kOffDO, p.dout + (static_cast<int64_t>(b) * p.T + t0) * p.lddo + h * kDh, p.lddo, kQ, tid, kBwdThreads);
  load_rows(sbase + L::kOffK, p.k + static_cast<int64_t>(b) * S * p.ldkv + h * kDh, p.ldkv, S, tid, kBwdThreads);
  load_rows(sbase + L::kOffV, p.v + static_cast<int64_t>(b) * S * p.ldkv + h * kDh, p.ldkv, S, tid, kBwdThreads);
  cp_async_commit();
  for (int j = tid; j < S; j += kBwdThreads) bias[j] = (p.keep == nullptr || p.keep[static_cast<int64_t>(b) * S + j]) ? 0.f : -INFINITY;
  cp_async_wait<0>();
  __syncthreads();
  if (tid < S / 16) {
    float acc = 0.f;
    for (int i = 0; i < 16; ++i) acc += bias[tid * 16 + i];
    smem[L::kOffBar + 24 + tid] = acc < 0.f;      // chunk of 16 keys holds a padded key
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t hi = umma_desc_hi_sw128(1024);

  if (warp == 0 && elect_one()) {
    // S = Q K^T -> columns [0, S);  dP = dO V^T -> columns [256, 256 + S)
    const uint32_t idesc = umma_idesc_bf16(kQ, S, false, false);
    const uint32_t q_lo = umma_desc_lo(sbase + L::kOffQ, 16), k_lo = umma_desc_lo(sbase + L::kOffK, 16);
    const uint32_t do_lo = umma_desc_lo(sbase + L::kOffDO, 16), v_lo = umma_desc_lo(sbase + L::kOffV, 16);
#pragma unroll
    for (int kk = 0; kk < kDh / 16; ++kk) umma_ss_lo(tmem, q_lo + kk * 2, k_lo + kk * 2, hi, idesc, kk > 0);
#pragma unroll
    for (int kk = 0; kk < kDh / 16; ++kk) umma_ss_lo(tmem + 256, do_lo + kk * 2, v_lo + kk * 2, hi, idesc, kk > 0);
    umma_commit(bars);
  }
  __syncwarp();
  const int rowi = tid & (kQ - 1), quarter = tid >> 7;
  const uint32_t row = static_cast<uint32_t>((b * p.H + h) * p.T + t0 + rowi);
  const float lse = p.lse[row], dl = p.delta[row];
  uint32_t seed_lo = 0, seed_hi = 0;
  const bool drop = p.seed != nullptr;
  if (drop) {
    const uint64_t sd = *p.seed;
    seed_lo = static_cast<uint32_t>(sd);
    seed_hi = static_cast<uint32_t>(sd >> 32);
  }
  bar_wait(bars, 0);
  tc_fence_after_sync();
  const uint32_t t_row = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const int qs = S >> 2;               // keys per thread (a multiple of 16)
  for (int c0 = quarter * qs; c0 < (quarter + 1) * qs; c0 += 16) {
    uint32_t r[16], d[16];
    tmem_ld16(t_row + c0, r);
    tmem_ld16(t_row + 256 + c0, d);
    tmem_ld_wait();
    uint32_t ds[8], pd[8];
    if (smem[L::kOffBar + 24 + (c0 >> 4)]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + bias[c0 + i]);
    }
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const float p0 = exp2f(fmaf(__uint_as_float(r[i]), p.c, -lse));
      const float p1 = exp2f(fmaf(__uint_as_float(r[i + 1]), p.c, -lse));
      float m0 = 1.f, m1 = 1.f;        // D = keep / (1 - p)
      if (drop) {
        const uint32_t hsh = drop_hash(seed_lo, seed_hi, row, static_cast<uint32_t>(c0 + i) >> 1, static_cast<uint32_t>(S) >> 1);
        m0 = (hsh & 0xffffu) < p.drop_thr ? 0.f : p.keep_scale;
        m1 = (hsh >> 16) < p.drop_thr ? 0.f : p.keep_scale;
      }
      const float g0 = p0 * (fmaf(m0, __uint_as_float(d[i]), -dl) * p.scale);
      const float g1 = p1 * (fmaf(m1, __uint_as_float(d[i + 1]), -dl) * p.scale);
      ds[i >> 1] = pack_bf16(g0, g1);
      pd[i >> 1] = pack_bf16(p0 * m0, p1 * m1);
    }
    // 16 keys = two 16-byte pieces of row `rowi` in the atom of these keys
    const uint32_t a_ds = sbase + L::kOffDS + (c0 >> 6) * L::kAtom, a_dp = sbase + L::kOffDP + (c0 >> 6) * L::kAtom;
    const int piece0 = (c0 & 63) >> 3;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      sts128(a_ds + sw128_offset(rowi, piece0 + i), make_uint4(ds[4 * i], ds[4 * i + 1], ds[4 * i + 2], ds[4 * i + 3]));
      sts128(a_dp + sw128_offset(rowi, piece0 + i), make_uint4(pd[4 * i], pd[4 * i + 1], pd[4 * i + 2], pd[4 * i + 3]));
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  const int nblk = (S + 127) >> 7;     // 128-key blocks (M of the dK / dV MMAs)
  if (warp == 0 && elect_one()) {
    tc_fence_after_sync();
    constexpr uint32_t idesc_kn = umma_idesc_bf16(kQ, kDh, false, true);     // A K-major, B MN-major
    constexpr uint32_t idesc_nn = umma_idesc_bf16(kQ, kDh, true, true);      // A MN-major, B MN-major
    const uint32_t k_n = umma_desc_lo(sbase + L::kOffK, 0), q_n = umma_desc_lo(sbase + L::kOffQ, 0), do_n = umma_desc_lo(sbase + L::kOffDO, 0);
    // dQ = dS K: contraction over the keys (dS K-major: 16 keys = 32 B inside an atom)
    for (int j = 0; j < S / 16; ++j)
      umma_ss_lo(tmem, umma_desc_lo(sbase + L::kOffDS + (j >> 2) * L::kAtom, 16) + (j & 3) * 2, k_n + j * (2048 >> 4), hi, idesc_kn, j > 0);
    // dK = dS^T Q, dV = (D o P)^T dO: contraction over the 128 queries (the same tiles read MN-major: M = keys, two 64-key atoms per block)
    for (int m = 0; m < nblk; ++m) {
      const uint32_t ds_n = umma_desc_lo(sbase + L::kOffDS + 2 * m * L::kAtom, L::kAtom), dp_n = umma_desc_lo(sbase + L::kOffDP + 2 * m * L::kAtom, L::kAtom);
#pragma unroll
      for (int j = 0; j < kQ / 16; ++j) umma_ss_lo(tmem + 64 + 64 * m, ds_n + j * (2048 >> 4), q_n + j * (2048 >> 4), hi, idesc_nn, j > 0);
#pragma unroll
      for (int j = 0; j < kQ / 16; ++j) umma_ss_lo(tmem + 192 + 64 * m, dp_n + j * (2048 >> 4), do_n + j * (2048 >> 4), hi, idesc_nn, j > 0);
    }
    umma_commit(bars + 1);
  }
  __syncwarp();
  bar_wait(bars + 1, 0);
  tc_fence_after_sync();
  if (quarter < 2) {
    // dQ row -> bf16 (a half row per thread)
    __nv_bfloat16* qrow = p.dq + (static_cast<int64_t>(b) * p.T + t0 + rowi) * p.lddq + h * kDh + quarter * 32;
    uint32_t r[32];
    tmem_ld32(t_row + quarter * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16(__uint_as_float(r[8 * i + 0]), __uint_as_float(r[8 * i + 1]));
      u.y = pack_bf16(__uint_as_float(r[8 * i + 2]), __uint_as_float(r[8 * i + 3]));
      u.z = pack_bf16(__uint_as_float(r[8 * i + 4]), __uint_as_float(r[8 * i + 5]));
      u.w = pack_bf16(__uint_as_float(r[8 * i + 6]), __uint_as_float(r[8 * i + 7]));
      reinterpret_cast<uint4*>(qrow)[i] = u;
    }
  }
  // quarter a of the threads takes accumulator a (dK block 0 | dK block 1 | dV block 0 | dV block 1): lane = key of the block; the tile's
  // partial sums are added to the fp32 gradient (CTAs that run together belong to different (batch, head) pairs: the query tile is the
  // slowest grid dimension, so the eight partial sums of an element do not queue up on one L2 address)
  {
    const int m = quarter & 1, is_v = quarter >> 1;
    const int key = m * 128 + rowi;
    if (m < nblk) {
      float* g = (is_v ? p.dv : p.dk) + (static_cast<int64_t>(b) * S + key) * p.lddkv + h * kDh;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t r[32];
        tmem_ld32(t_row + (is_v ? 192 : 64) + 64 * m + hh * 32, r);
        tmem_ld_wait();
#ifdef XA_NORED
        if (key < -1) {
#else
        if (key < S) {
#endif
#pragma unroll
          for (int i = 0; i < 8; ++i)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g + hh * 32 + 4 * i), "f"(__uint_as_float(r[4 * i])), "f"(__uint_as_float(r[4 * i + 1])),
                         "f"(__uint_as_float(r[4 * i + 2])), "f"(__uint_as_float(r[4 * i + 3]))
                         : "memory");
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace xa
}  // namespace rtts

using namespace rtts;

static long long* g_xa_trace = nullptr;
extern "C" void rtts_debug_set_xattn_trace(void* device_buffer) { g_xa_trace = static_cast<long long*>(device_buffer); }

static int xattn_check(const char* fn, int B, int T, int S, int H, int dh) {
  if (dh != xa::kDh) return fail(kErrBadArg, "%s: head size %d unsupported (64 only)", fn, dh);
  if (B <= 0 || H <= 0 || T <= 0 || T % xa::kQ != 0) return fail(kErrBadArg, "%s: T=%d must be a positive multiple of 128", fn, T);
  if (S <= 0 || S > xa::kMaxS || S % 32 != 0) return fail(kErrBadArg, "%s: S=%d must be a multiple of 32 in [32, 256]", fn, S);
  return 0;
}

extern "C" int rtts_xattn_fwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const uint8_t* keep, float scale, float p_drop,
                              const uint64_t* seed, void* out, int64_t ldo, float* lse, int B, int T, int S, int H, int dh, void* stream) {
  RTTS_REQUIRE(q && k && v && out && lse, "rtts_xattn_fwd: null pointer");
  if (int e = xattn_check("rtts_xattn_fwd", B, T, S, H, dh)) return e;
  RTTS_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0 &&
                   ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
               "rtts_xattn_fwd: tensors must be 16-byte aligned");
  RTTS_REQUIRE(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || seed), "rtts_xattn_fwd: dropout needs a seed word and 0 <= p < 1");
  xa::Params p = {};
  p.q = static_cast<const __nv_bfloat16*>(q); p.ldq = ldq;
  p.k = static_cast<const __nv_bfloat16*>(k); p.v = static_cast<const __nv_bfloat16*>(v); p.ldkv = ldkv;
  p.keep = keep;
  p.out = static_cast<__nv_bfloat16*>(out); p.ldo = ldo;
  p.lse = lse;
  p.seed = p_drop > 0.f ? seed : nullptr;
  p.drop_thr = static_cast<uint32_t>(p_drop * 65536.f);
  p.keep_scale = 1.f / (1.f - p_drop);
  p.scale = scale;
  p.c = scale * xa::kLog2e;
  p.B = B; p.T = T; p.S = S; p.H = H;
  p.trace = g_xa_trace;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(xa::xattn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, xa::FwdSmem::kTotal);
    if (e != cudaSuccess) return fail(kErrCuda, "rtts_xattn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  // (the query tile is the slowest grid dimension: CTAs that run together then belong to different (batch, head) pairs)
  xa::xattn_fwd_kernel<<<dim3(H, B, T / xa::kQ), xa::kFwdThreads, xa::FwdSmem::kTotal, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("rtts_xattn_fwd");
}

extern "C" int rtts_xattn_bwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const uint8_t* keep, float scale, float p_drop,
                              const uint64_t* seed, const void* dout, int64_t lddo, const float* lse, const float* delta, void* dq, int64_t lddq,
                              float* dk, float* dv, int64_t lddkv, int B, int T, int S, int H, int dh, void* stream) {
  RTTS_REQUIRE(q && k && v && dout && lse && delta && dq && dk && dv, "rtts_xattn_bwd: null pointer");
  if (int e = xattn_check("rtts_xattn_bwd", B, T, S, H, dh)) return e;
  RTTS_REQUIRE(S % 64 == 0, "rtts_xattn_bwd: S=%d must be a multiple of 64", S);
  RTTS_REQUIRE(ldq % 8 == 0 && ldkv % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 && lddkv % 4 == 0 &&
                   ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(dout) |
                     reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 15) == 0,
               "rtts_xattn_bwd: tensors must be 16-byte aligned");
  RTTS_REQUIRE(p_drop >= 0.f && p_drop < 1.f && (p_drop == 0.f || seed), "rtts_xattn_bwd: dropout needs a seed word and 0 <= p < 1");
  xa::Params p = {};
  p.q = static_cast<const __nv_bfloat16*>(q); p.ldq = ldq;
  p.k = static_cast<const __nv_bfloat16*>(k); p.v = static_cast<const __nv_bfloat16*>(v); p.ldkv = ldkv;
  p.keep = keep;
  p.dout = static_cast<const __nv_bfloat16*>(dout); p.lddo = lddo;
  p.lse = const_cast<float*>(lse); p.delta = delta;
  p.dq = static_cast<__nv_bfloat16*>(dq); p.lddq = lddq;
  p.dk = dk; p.dv = dv; p.lddkv = lddkv;
  p.seed = p_drop > 0.f ? seed : nullptr;
  p.drop_thr = static_cast<uint32_t>(p_drop * 65536.f);
  p.keep_scale = 1.f / (1.f - p_drop);
  p.scale = scale;
  p.c = scale * xa::kLog2e;
  p.B = B; p.T = T; p.S = S; p.H = H;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(xa::xattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, xa::BwdSmem::kTotal);
    if (e != cudaSuccess) return fail(kErrCuda, "rtts_xattn_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  xa::xattn_bwd_kernel<<<dim3(H, B, T / xa::kQ), xa::kBwdThreads, xa::BwdSmem::kTotal, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("rtts_xattn_bwd");
}
