// Fused chunked shared-QK attention, forward (rtts_lsh_attn_fwd) and the round merge
// (rtts_lsh_merge_fwd).  See include/rtts_b200.h for the contract and DESIGN.md "K9" for the tiling.
//
// One CTA (128 threads) owns 128 consecutive SORTED query slots of one (batch, head) row:
//   bucket 64 : two chunks c0,c1; key tile = slots of chunks [c0-1, c0, c1]      (192 rows)
//   bucket 128: one chunk c;      key tile = slots of chunks [c-1, c]            (256 rows)
// and the queries are simply the last 128 rows of the key tile (shared-QK: one gather serves both
// operands).  Rows are gathered straight from the UNSORTED token-major qk / v arrays through
// `sticker` with 16-byte cp.async into SWIZZLE_128B shared-memory tiles (a row is 64 bf16 = one
// 128-byte swizzle row), so the R sorted copies the reference materialises never exist.
//   S = Q K^T          tcgen05.mma  M=128, N=192|256, K=64   -> TMEM fp32
//   softmax            one thread per query row reads its 2*bucket-wide window from TMEM, applies the
//                      per-key 1/|k| scale (keys are normalised AFTER the fp32-accumulated dot), the
//                      padding / causal / self masks from the position ids, two passes (max, exp)
//   O = P V            P (bf16) goes to shared memory in K-major SW128 layout, V is the gathered tile
//                      used MN-major; tcgen05.mma M=128, N=64, K=192|256 -> TMEM (aliases S)
//   epilogue           O / rowsum -> bf16, scatter-stored at the UNSORTED slot (r*T + pos); lse too.
#include <cfloat>

#include "common.cuh"
#include "host_util.h"
#include "rtts_b200.h"

namespace rtts {

constexpr int kDh = 64;
constexpr int kQRows = 128;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kPadFlag = 0x40000000;

struct AttnFwdParams {
  const __nv_bfloat16* qk;
  const __nv_bfloat16* v;
  int64_t ld;
  const int32_t* sticker;
  const uint8_t* mask;
  __nv_bfloat16* o_rounds;
  float* lse_rounds;
  int T, H, R;
  int tiles_per_row;  // R*T / 128
  float score_scale_log2;  // score_scale * log2(e)
  float mask_value_log2, self_value_log2;
  int key_norm, mask_mode, causal;
};

template <int BUCKET>
struct AttnFwdSmem {
  static constexpr int kKeyRows = kQRows + BUCKET;
  static constexpr int kKeyBytes = kKeyRows * 128;
  static constexpr int kPBytes = kQRows * kKeyRows * 2;   // P tile, bf16
  // layout (all tile bases 1024-B aligned):  [ K tile | V tile | P tile | small arrays ]
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kKeyBytes;
  static constexpr int kOffP = 2 * kKeyBytes;
  static constexpr int kOffScale = kOffP + kPBytes;                 // float[kKeyRows]
  static constexpr int kOffPos = kOffScale + kKeyRows * 4;          // int[kKeyRows]
  static constexpr int kOffSlot = kOffPos + kKeyRows * 4;           // int[kQRows]  unsorted slot of each query
  static constexpr int kOffBar = kOffSlot + kQRows * 4;             // uint64 mbarrier
  static constexpr int kOffTmem = kOffBar + 8;                      // uint32
  static constexpr int kTotal = kOffTmem + 8;
  static constexpr int kDynamic = kTotal + 1024;                    // slack for manual 1024-B alignment
};

template <int BUCKET>
__global__ void __launch_bounds__(128) lsh_attn_fwd_kernel(const AttnFwdParams p) {
  using L = AttnFwdSmem<BUCKET>;
  constexpr int kKeyRows = L::kKeyRows;
  constexpr int kQOff = BUCKET;           // first query row inside the key tile
  constexpr int kWin = 2 * BUCKET;        // attention window per query
  constexpr uint32_t kTmemCols = 256;     // S uses kKeyRows (<=256) columns; O aliases S[0..64)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sK = smem_u32(smem + L::kOffK), sV = smem_u32(smem + L::kOffV), sP = smem_u32(smem + L::kOffP);
  float* key_scale = reinterpret_cast<float*>(smem + L::kOffScale);
  int* key_pos = reinterpret_cast<int*>(smem + L::kOffPos);
  int* q_slot = reinterpret_cast<int*>(smem + L::kOffSlot);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmem);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row_bh = blockIdx.x / p.tiles_per_row;           // (batch*H + head)
  const int tile = blockIdx.x - row_bh * p.tiles_per_row;
  const int b = row_bh / p.H, h = row_bh - b * p.H;
  const int RT = p.R * p.T;
  const int32_t* stk = p.sticker + static_cast<int64_t>(row_bh) * RT;
  const int first_slot = tile * kQRows - BUCKET;               // sorted slot of key-tile row 0 (may be < 0: wraps)

  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }

  // ---- gather: 8 lanes per row, 16 rows per pass -------------------------------------------------
  // sticker[slot] = round*T + pos and sorted slots keep the rounds contiguous, so pos = sticker - (slot / T) * T:
  // the round comes from the slot index, not from the loaded value (no integer division on the load's critical path).
  {
    const int g = tid >> 3, c = tid & 7;
    constexpr int kPasses = kKeyRows / 16;
    const int prev_first = first_slot < 0 ? first_slot + RT : first_slot;       // slot of key-tile row 0 (look-back chunk)
    const int base_prev = (prev_first / p.T) * p.T, base_main = ((tile * kQRows) / p.T) * p.T;
    int st[kPasses];
#pragma unroll
    for (int i = 0; i < kPasses; ++i) {
      const int j = i * 16 + g;
      st[i] = __ldg(stk + (j < BUCKET ? prev_first + j : first_slot + j));
    }
    int pos[kPasses];
    uint8_t valid[kPasses];
#pragma unroll
    for (int i = 0; i < kPasses; ++i) {
      pos[i] = st[i] - ((i * 16 + g) < BUCKET ? base_prev : base_main);
      valid[i] = (p.mask != nullptr && c == 0) ? __ldg(p.mask + static_cast<int64_t>(b) * p.T + pos[i]) : uint8_t(1);
    }
    const int64_t head_off = static_cast<int64_t>(h) * kDh + c * 8;
#pragma unroll
    for (int i = 0; i < kPasses; ++i) {
      const int j = i * 16 + g;
      const int64_t off = (static_cast<int64_t>(b) * p.T + pos[i]) * p.ld + head_off;
      const uint32_t so = sw128_offset(j, c);
      cp_async16(sK + so, p.qk + off);
      cp_async16(sV + so, p.v + off);
    }
    cp_async_commit();
    if (c == 0) {
#pragma unroll
      for (int i = 0; i < kPasses; ++i) {
        const int j = i * 16 + g;
        key_pos[j] = valid[i] ? pos[i] : (pos[i] | kPadFlag);
        if (j >= kQOff) q_slot[j - kQOff] = st[i];
      }
    }
    cp_async_wait<0>();
  }
  __syncthreads();

  // ---- per-key scale: 1/|k| (or rms variant) * score_scale * log2(e) --------------------------------
  for (int j = tid; j < kKeyRows; j += 128) {
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 u = *reinterpret_cast<const uint4*>(smem + L::kOffK + sw128_offset(j, c));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float lo = bf16_lo(w[e]), hi = bf16_hi(w[e]);
        ss = fmaf(lo, lo, ss);
        ss = fmaf(hi, hi, ss);
      }
    }
    float inv;
    if (p.key_norm == RTTS_KEYNORM_L2) inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    else inv = rsqrtf(ss * (1.f / kDh) + 1e-6f) * 0.125f;   // 1/sqrt(64)
    key_scale[j] = inv * p.score_scale_log2;
  }
  fence_proxy_async_smem();   // cp.async / st.shared data -> visible to the tensor-core (async) proxy
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  // ---- S = Q K^T -------------------------------------------------------------------------------------
  if (tid == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, kKeyRows, false, false);
#pragma unroll
    for (int k = 0; k < kDh / 16; ++k) {
      const uint64_t da = umma_desc_sw128(sK + kQOff * 128 + k * 32, 16, 1024);
      const uint64_t db = umma_desc_sw128(sK + k * 32, 16, 1024);
      umma_ss(tmem, da, db, idesc, k > 0);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after_sync();

  // ---- softmax over this thread's query row ----------------------------------------------------------
  const int m = tid;
  const int q_enc = key_pos[kQOff + m];
  int q_limit = p.causal ? (q_enc & ~kPadFlag) : (kPadFlag - 1);
  if (p.mask_mode == RTTS_MASK_QUERY_AND_KEY && (q_enc & kPadFlag)) q_limit = -1;   // padded query: all masked
  const int win0 = (m / BUCKET) * BUCKET;                      // first key-tile row of this query's window
  const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const float mv = p.mask_value_log2, sv = p.self_value_log2;

  float row_max = -FLT_MAX;
#pragma unroll 1
  for (int c0 = 0; c0 < kWin; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(t_row + win0 + c0, r);
    tmem_ld_wait();
    int kp[32];
    float ks[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {      // win0 + c0 is a multiple of 32: 16-byte aligned broadcast loads
      *reinterpret_cast<int4*>(kp + 4 * i) = *reinterpret_cast<const int4*>(key_pos + win0 + c0 + 4 * i);
      *reinterpret_cast<float4*>(ks + 4 * i) = *reinterpret_cast<const float4*>(key_scale + win0 + c0 + 4 * i);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float s = __uint_as_float(r[i]) * ks[i];
      s = kp[i] > q_limit ? mv : s;
      s = kp[i] == q_enc ? sv : s;
      row_max = fmaxf(row_max, s);
    }
  }
  float row_sum = 0.f;
  uint8_t* p_row_base = smem + L::kOffP;
#pragma unroll 1
  for (int c0 = 0; c0 < kWin; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(t_row + win0 + c0, r);
    tmem_ld_wait();
    float e[32];
    int kp[32];
    float ks[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      *reinterpret_cast<int4*>(kp + 4 * i) = *reinterpret_cast<const int4*>(key_pos + win0 + c0 + 4 * i);
      *reinterpret_cast<float4*>(ks + 4 * i) = *reinterpret_cast<const float4*>(key_scale + win0 + c0 + 4 * i);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float s = __uint_as_float(r[i]) * ks[i];
      s = kp[i] > q_limit ? mv : s;
      s = kp[i] == q_enc ? sv : s;
      e[i] = exp2f(s - row_max);
      row_sum += e[i];
    }
    // keys win0+c0 .. +31 -> P tile k-block (col/64), 16-byte chunk (col%64)/8
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const int col = win0 + c0 + q4 * 8;
      uint4 u;
      u.x = pack_bf16(e[q4 * 8 + 0], e[q4 * 8 + 1]);
      u.y = pack_bf16(e[q4 * 8 + 2], e[q4 * 8 + 3]);
      u.z = pack_bf16(e[q4 * 8 + 4], e[q4 * 8 + 5]);
      u.w = pack_bf16(e[q4 * 8 + 6], e[q4 * 8 + 7]);
      *reinterpret_cast<uint4*>(p_row_base + (col >> 6) * (kQRows * 128) + sw128_offset(m, (col & 63) >> 3)) = u;
    }
  }
  if (BUCKET == 64) {
    // the 64 key rows outside this query's window contribute nothing: zero that k-block of P
    const int dead_block = (m < 64) ? 2 : 0;
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int c = 0; c < 8; ++c)
      *reinterpret_cast<uint4*>(p_row_base + dead_block * (kQRows * 128) + sw128_offset(m, c)) = z;
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();          // every row of S has been consumed; P is complete
  tc_fence_after_sync();

  // ---- O = P V  (O aliases the first 64 columns of S) ---------------------------------------------------
  if (tid == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, kDh, false, true);
#pragma unroll
    for (int j = 0; j < kKeyRows / 16; ++j) {
      const uint64_t da = umma_desc_sw128(sP + (j >> 2) * (kQRows * 128) + (j & 3) * 32, 16, 1024);
      const uint64_t db = umma_desc_sw128(sV + j * 2048, 0, 1024);
      umma_ss(tmem, da, db, idesc, j > 0);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 1);
  tc_fence_after_sync();

  // ---- epilogue -------------------------------------------------------------------------------------------
  {
    const float inv_sum = 1.f / row_sum;
    const int64_t slot = static_cast<int64_t>(row_bh) * RT + q_slot[m];
    uint4* dst = reinterpret_cast<uint4*>(p.o_rounds + slot * kDh);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t r[32];
      tmem_ld32(t_row + half * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 u;
        u.x = pack_bf16(__uint_as_float(r[q4 * 8 + 0]) * inv_sum, __uint_as_float(r[q4 * 8 + 1]) * inv_sum);
        u.y = pack_bf16(__uint_as_float(r[q4 * 8 + 2]) * inv_sum, __uint_as_float(r[q4 * 8 + 3]) * inv_sum);
        u.z = pack_bf16(__uint_as_float(r[q4 * 8 + 4]) * inv_sum, __uint_as_float(r[q4 * 8 + 5]) * inv_sum);
        u.w = pack_bf16(__uint_as_float(r[q4 * 8 + 6]) * inv_sum, __uint_as_float(r[q4 * 8 + 7]) * inv_sum);
        dst[half * 4 + q4] = u;
      }
    }
    p.lse_rounds[slot] = (row_max + log2f(row_sum)) * kLn2;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------------------
// Round merge (rp R11): out[t] = sum_r o_r[t] * exp(lse_r[t] - logsumexp_r lse_r[t]).  8 lanes per (b,h,t) row,
// each lane owns 16 bytes of the 128-byte head slice.  HBM-bound: reads R*(128+4) B, writes 128+4 B per row.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lsh_merge_fwd_kernel(const __nv_bfloat16* __restrict__ o_rounds,
                                                            const float* __restrict__ lse_rounds,
                                                            __nv_bfloat16* __restrict__ out, int64_t ld_out,
                                                            float* __restrict__ lse, int T, int H, int R, int64_t rows) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x) >> 3;   // (b*H + h)*T + t
  const int c = threadIdx.x & 7;
  if (row >= rows) return;
  const int64_t bh = row / T;
  const int t = static_cast<int>(row - bh * T);
  const float* l = lse_rounds + bh * R * T + t;
  float mx = -FLT_MAX;
  for (int r = 0; r < R; ++r) mx = fmaxf(mx, l[static_cast<int64_t>(r) * T]);
  float den = 0.f;
  for (int r = 0; r < R; ++r) den += __expf(l[static_cast<int64_t>(r) * T] - mx);
  const float inv_den = 1.f / den;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int r = 0; r < R; ++r) {
    const float w = __expf(l[static_cast<int64_t>(r) * T] - mx) * inv_den;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(o_rounds + ((bh * R + r) * T + t) * kDh) + c);
    acc[0] = fmaf(w, bf16_lo(u.x), acc[0]); acc[1] = fmaf(w, bf16_hi(u.x), acc[1]);
    acc[2] = fmaf(w, bf16_lo(u.y), acc[2]); acc[3] = fmaf(w, bf16_hi(u.y), acc[3]);
    acc[4] = fmaf(w, bf16_lo(u.z), acc[4]); acc[5] = fmaf(w, bf16_hi(u.z), acc[5]);
    acc[6] = fmaf(w, bf16_lo(u.w), acc[6]); acc[7] = fmaf(w, bf16_hi(u.w), acc[7]);
  }
  const int64_t b = bh / H;
  const int h = static_cast<int>(bh - b * H);
  uint4 u;
  u.x = pack_bf16(acc[0], acc[1]); u.y = pack_bf16(acc[2], acc[3]);
  u.z = pack_bf16(acc[4], acc[5]); u.w = pack_bf16(acc[6], acc[7]);
  reinterpret_cast<uint4*>(out + (b * T + t) * ld_out + h * kDh)[c] = u;
  if (c == 0) lse[row] = mx + __logf(den);
}

template <int BUCKET>
int launch_attn_fwd(const AttnFwdParams& p, int ctas, cudaStream_t stream) {
  using L = AttnFwdSmem<BUCKET>;
  static bool configured = false;   // idempotent attribute set; benign if raced
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(lsh_attn_fwd_kernel<BUCKET>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
    if (e != cudaSuccess) return fail(kErrCuda, "rtts_lsh_attn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  lsh_attn_fwd_kernel<BUCKET><<<ctas, 128, L::kDynamic, stream>>>(p);
  return check_launch("rtts_lsh_attn_fwd");
}

}  // namespace rtts

using namespace rtts;

extern "C" int rtts_lsh_attn_fwd(const void* qk, const void* v, int64_t ld, const int32_t* sticker, const uint8_t* mask,
                                 const rtts_lsh_spec* spec, void* o_rounds, float* lse_rounds, int B, int T, int H, int dh,
                                 int R, int bucket, void* stream) {
  RTTS_REQUIRE(qk && v && sticker && spec && o_rounds && lse_rounds, "rtts_lsh_attn_fwd: null pointer");
  RTTS_REQUIRE(dh == kDh, "rtts_lsh_attn_fwd: head size %d unsupported (64 only)", dh);
  RTTS_REQUIRE(bucket == 64 || bucket == 128, "rtts_lsh_attn_fwd: bucket size %d unsupported (64 or 128)", bucket);
  RTTS_REQUIRE(T % (2 * bucket) == 0, "rtts_lsh_attn_fwd: T=%d must be a multiple of 2*bucket", T);
  RTTS_REQUIRE(ld % 8 == 0 && ((reinterpret_cast<uintptr_t>(qk) | reinterpret_cast<uintptr_t>(v) |
                                reinterpret_cast<uintptr_t>(o_rounds)) & 15) == 0,
               "rtts_lsh_attn_fwd: tensors must be 16-byte aligned");
  RTTS_REQUIRE(static_cast<int64_t>(T) < kPadFlag, "rtts_lsh_attn_fwd: T too large");
  AttnFwdParams p;
  p.qk = static_cast<const __nv_bfloat16*>(qk);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.ld = ld;
  p.sticker = sticker;
  p.mask = mask;
  p.o_rounds = static_cast<__nv_bfloat16*>(o_rounds);
  p.lse_rounds = lse_rounds;
  p.T = T; p.H = H; p.R = R;
  p.tiles_per_row = R * T / kQRows;
  p.score_scale_log2 = spec->score_scale * kLog2e;
  p.mask_value_log2 = fmaxf(spec->mask_value * kLog2e, -3.0e38f);
  p.self_value_log2 = spec->self_value * kLog2e;
  p.key_norm = spec->key_norm; p.mask_mode = spec->mask_mode; p.causal = spec->causal;
  const int64_t ctas = static_cast<int64_t>(B) * H * p.tiles_per_row;
  RTTS_REQUIRE(ctas > 0 && ctas < (1ll << 31), "rtts_lsh_attn_fwd: bad grid");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return bucket == 64 ? launch_attn_fwd<64>(p, static_cast<int>(ctas), s) : launch_attn_fwd<128>(p, static_cast<int>(ctas), s);
}

extern "C" int rtts_lsh_merge_fwd(const void* o_rounds, const float* lse_rounds, void* out, int64_t ld_out, float* lse, int B,
                                  int T, int H, int dh, int R, void* stream) {
  RTTS_REQUIRE(o_rounds && lse_rounds && out && lse, "rtts_lsh_merge_fwd: null pointer");
  RTTS_REQUIRE(dh == kDh, "rtts_lsh_merge_fwd: head size %d unsupported (64 only)", dh);
  RTTS_REQUIRE(ld_out % 8 == 0, "rtts_lsh_merge_fwd: ld_out must be a multiple of 8");
  const int64_t rows = static_cast<int64_t>(B) * H * T;
  const int64_t blocks = (rows * 8 + 255) / 256;
  lsh_merge_fwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(o_rounds), lse_rounds, static_cast<__nv_bfloat16*>(out), ld_out, lse, T, H, R, rows);
  return check_launch("rtts_lsh_merge_fwd");
}
