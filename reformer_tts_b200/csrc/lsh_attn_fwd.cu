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
  const float* sumsq;   // [B,H,T] |qk row|^2
  const uint8_t* mask;
  __nv_bfloat16* o_rounds;
  float* lse_rounds;
  long long* trace;     // debug: per-role clock64 stamps of CTA 0 (nullable)
  int T, H, R;
  int tiles_per_row;  // R*T / 128
  float score_scale_log2;  // score_scale * log2(e)
  float mask_value_log2, self_value_log2;
  int key_norm, mask_mode, causal;
};

// Persistent, warp-specialised pipeline: one CTA per SM walks tiles t = blockIdx.x, += gridDim.x.
//   warp 12     : TMEM allocation; lane 0 issues every tcgen05.mma  (S(n) as soon as K lands, then PV(n-1) when P(n-1) is ready)
//   warps 8-11  : loaders - sticker -> position -> qk and V rows by 16-byte cp.async into swizzled tiles, key scale from sumsq;
//                 they run up to kStages tiles ahead of the consumers
//   warps 0-3   : softmax group 0 (tiles n even);  warps 4-7: softmax group 1 (tiles n odd); thread = query row = TMEM lane
// Shared memory: kStages stage sets {K|P tile, V tile, key scale / position / slot arrays}; P aliases the K tile of its stage
// (K is dead once S(n) has completed, which s_full certifies).  TMEM: one 256-column region per softmax group: S in
// [0, kKeyRows), O in [192, 256) for bucket 64 (disjoint) or aliased onto S[0, 64) for bucket 128.
// mbarriers: kv_full[stage] (128 loader arrivals) -> s_full[group] (tcgen05.commit) -> p_full[group] (128 softmax arrivals)
//            -> o_full[group] + kv_free[stage] (tcgen05.commit after PV) -> o_free[group] (128 arrivals after the epilogue).
//
// Softmax is single-pass: keys are unit vectors after normalisation, so |s_ij| <= |q_i| * score_scale (Cauchy-Schwarz) and
// m_i = that bound is a valid stabiliser known before any score is read; masked / self entries give exp2() == 0 exactly.  If a
// row sums to zero (its only visible key is itself - masked to self_value - or, for absurd norms, everything underflowed) the
// warp re-does that row with the exact two-pass arithmetic of the reference (row max first), so results never depend on the bound.
constexpr int kFwdThreads = 416;
constexpr int kLoaderThreads = 128;
// Warp roles by index.  The SM's issue arbiter favours higher warp ids, so the producers that everything else waits on get
// the top ids: warps 0-7 softmax groups, warps 8-11 loaders, warp 12 MMA issuer.
constexpr int kFirstLoaderWarp = 8;
constexpr int kMmaWarp = 12;

template <int BUCKET>
struct AttnFwdSmem {
  static constexpr int kKeyRows = kQRows + BUCKET;
  static constexpr int kStages = BUCKET == 64 ? 3 : 2;
  static constexpr int kKeyBytes = kKeyRows * 128;
  static constexpr int kPBytes = kQRows * kKeyRows * 2;   // P tile, bf16 (>= kKeyBytes): shares its storage with the K tile
  // per stage (all tile bases 1024-B aligned):  [ K|P tile | V tile | scale | pos | slot ]
  static constexpr int kOffKP = 0;
  static constexpr int kOffV = kPBytes;
  static constexpr int kOffScale = kOffV + kKeyBytes;                // float[kKeyRows]
  static constexpr int kOffPos = kOffScale + kKeyRows * 4;           // int[kKeyRows]
  static constexpr int kOffSlot = kOffPos + kKeyRows * 4;            // int[kQRows]  unsorted slot of each query
  static constexpr int kStageBytes = ((kOffSlot + kQRows * 4 + 1023) / 1024) * 1024;
  static constexpr int kOffBar = kStages * kStageBytes;              // 2*kStages + 8 mbarriers
  static constexpr int kOffTmem = kOffBar + (2 * kStages + 8) * 8;
  static constexpr int kTotal = kOffTmem + 8;
  static constexpr int kDynamic = kTotal + 1024;                     // slack for manual 1024-B alignment
};

#define RTTS_STAMP(role, n, k) do { if (p.trace != nullptr && blockIdx.x == 0 && (n) < 32) p.trace[((role) * 32 + (n)) * 8 + (k)] = clock64(); } while (0)

// One 32-column chunk of the single-pass softmax for one query row: scores r -> e = exp2(s * key_scale - bound), zero where the
// key is masked (MASK) or is the query itself (SELF), accumulate the row sum, store the bf16 P chunk (4 x 16 B, swizzled).
// a_pos / a_scale: shared addresses of key_pos / key_scale at this chunk's first column; a_p: shared address of P row m, k-block
// of this chunk; c16: index of the chunk's first 16-byte column group inside that k-block; m7 = m & 7 (swizzle phase).
template <bool MASK, bool SELF>
__device__ __forceinline__ void soft_chunk(const uint32_t* r, uint32_t a_pos, uint32_t a_scale, uint32_t a_p, int c16, int m7, float neg_bound,
                                           int q_limit, int q_enc, float* sum4) {
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    const uint4 s0 = lds128(a_scale + q4 * 32), s1 = lds128(a_scale + q4 * 32 + 16);
    const float ks[8] = {__uint_as_float(s0.x), __uint_as_float(s0.y), __uint_as_float(s0.z), __uint_as_float(s0.w),
                         __uint_as_float(s1.x), __uint_as_float(s1.y), __uint_as_float(s1.z), __uint_as_float(s1.w)};
    float e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = exp2f(fmaf(__uint_as_float(r[q4 * 8 + i]), ks[i], neg_bound));
    if (MASK || SELF) {
      const uint4 p0 = lds128(a_pos + q4 * 32), p1 = lds128(a_pos + q4 * 32 + 16);
      const int kp[8] = {static_cast<int>(p0.x), static_cast<int>(p0.y), static_cast<int>(p0.z), static_cast<int>(p0.w),
                         static_cast<int>(p1.x), static_cast<int>(p1.y), static_cast<int>(p1.z), static_cast<int>(p1.w)};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MASK) e[i] = kp[i] > q_limit ? 0.f : e[i];       // exp2(mask_value - m) == 0
        if (SELF) e[i] = kp[i] == q_enc ? 0.f : e[i];        // exp2(self_value - m) == 0
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sum4[i & 3] += e[i];
    uint4 u;
    u.x = pack_bf16(e[0], e[1]);
    u.y = pack_bf16(e[2], e[3]);
    u.z = pack_bf16(e[4], e[5]);
    u.w = pack_bf16(e[6], e[7]);
    sts128(a_p + (((c16 + q4) ^ m7) << 4), u);
  }
}

template <int BUCKET>
__global__ void __launch_bounds__(kFwdThreads, 1) lsh_attn_fwd_kernel(const AttnFwdParams p, const int num_tiles) {
  using L = AttnFwdSmem<BUCKET>;
  constexpr int kKeyRows = L::kKeyRows, kStages = L::kStages;
  constexpr int kQOff = BUCKET;           // first query row inside the key tile
  constexpr int kWin = 2 * BUCKET;        // attention window per query
  constexpr bool kAliasO = BUCKET == 128; // bucket 128: S fills all 256 columns of the region, O reuses S[0, 64)
  constexpr uint32_t kColO = kAliasO ? 0 : 192;
  constexpr uint32_t kTmemCols = 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* kv_full = bars;                    // [kStages]
  uint64_t* kv_free = bars + kStages;          // [kStages]
  uint64_t* s_full = bars + 2 * kStages;       // [2]
  uint64_t* p_full = s_full + 2;               // [2]
  uint64_t* o_full = s_full + 4;               // [2]
  uint64_t* o_free = s_full + 6;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int RT = p.R * p.T;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(kv_full + s, kLoaderThreads);
      mbar_init(kv_free + s, 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full + g, 1);
      mbar_init(p_full + g, kQRows);
      mbar_init(o_full + g, 1);
      mbar_init(o_free + g, kQRows);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == kMmaWarp) {
    // ================================================= MMA issuer =================================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kKeyRows, false, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kDh, false, true);
      auto issue_pv = [&](int m) {      // O(m) = P(m) V(m)
        const int g = m & 1, st = m % kStages;
        const uint32_t sP = smem_u32(smem + st * L::kStageBytes + L::kOffKP), sV = smem_u32(smem + st * L::kStageBytes + L::kOffV);
        mbar_wait(p_full + g, (m >> 1) & 1);
        RTTS_STAMP(0, m, 2);
        if (!kAliasO) mbar_wait(o_free + g, ((m >> 1) & 1) ^ 1);      // epilogue of tile m-2 has drained O of this group
        tc_fence_after_sync();
#pragma unroll
        for (int j = 0; j < kKeyRows / 16; ++j)
          umma_ss(tmem + g * 256 + kColO, umma_desc_sw128(sP + (j >> 2) * (kQRows * 128) + (j & 3) * 32, 16, 1024),
                  umma_desc_sw128(sV + j * 2048, 0, 1024), idesc_o, j > 0);
        umma_commit(o_full + g);
        umma_commit(kv_free + st);
        RTTS_STAMP(0, m, 3);      // the stage's K|P and V tiles are no longer read
      };
      int n = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n) {
        const int g = n & 1, st = n % kStages;
        const uint32_t sKP = smem_u32(smem + st * L::kStageBytes + L::kOffKP);
        mbar_wait(kv_full + st, (n / kStages) & 1);
        RTTS_STAMP(0, n, 0);
        if (kAliasO) mbar_wait(o_free + g, ((n >> 1) & 1) ^ 1);       // S(n) overwrites O(n-2): its epilogue must be done
        tc_fence_after_sync();
        // S region of group g is free: PV(n-2) was issued (program order) after p_full(n-2), i.e. after the last read of S(n-2)
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k)
          umma_ss(tmem + g * 256, umma_desc_sw128(sKP + kQOff * 128 + k * 32, 16, 1024), umma_desc_sw128(sKP + k * 32, 16, 1024), idesc_s, k > 0);
        umma_commit(s_full + g);
        RTTS_STAMP(0, n, 1);
        if (n > 0) issue_pv(n - 1);
      }
      if (n > 0) issue_pv(n - 1);
    }
  } else if (warp >= kFirstLoaderWarp && warp < kMmaWarp) {
    // ================================================= loaders ====================================================
    // sticker[slot] = round*T + pos and sorted slots keep the rounds contiguous, so pos = sticker - (slot / T) * T with the
    // round taken from the slot index (no integer division on the load's critical path).
    const int lt = tid - kFirstLoaderWarp * 32;  // 0..127
    const int grp = lt >> 3, c = lt & 7;         // 16 row groups of 8 lanes; lane c owns 16-byte chunk c of a row
    constexpr int kPasses = kKeyRows / 16;
    int n = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++n) {
      const int st_i = n % kStages;
      uint8_t* stage = smem + st_i * L::kStageBytes;
      const uint32_t sKP = smem_u32(stage + L::kOffKP), sV = smem_u32(stage + L::kOffV);
      float* key_scale = reinterpret_cast<float*>(stage + L::kOffScale);
      int* key_pos = reinterpret_cast<int*>(stage + L::kOffPos);
      int* q_slot = reinterpret_cast<int*>(stage + L::kOffSlot);
      const int row_bh = tile / p.tiles_per_row, t_in = tile - row_bh * p.tiles_per_row;
      const int b = row_bh / p.H, h = row_bh - b * p.H;
      const int32_t* stk = p.sticker + static_cast<int64_t>(row_bh) * RT;
      const int first_slot = t_in * kQRows - BUCKET;
      const int prev_first = first_slot < 0 ? first_slot + RT : first_slot;
      const int base_prev = (prev_first / p.T) * p.T, base_main = ((t_in * kQRows) / p.T) * p.T;
      if (lt == 0) RTTS_STAMP(1, n, 0);
      int st[kPasses];
#pragma unroll
      for (int i = 0; i < kPasses; ++i) {
        const int j = i * 16 + grp;
        st[i] = __ldg(stk + (j < BUCKET ? prev_first + j : first_slot + j));
      }
      float ssq[kPasses];
      uint8_t valid[kPasses];
      if (c == 0) {
        const float* sq = p.sumsq + static_cast<int64_t>(row_bh) * p.T;
#pragma unroll
        for (int i = 0; i < kPasses; ++i) {
          const int pos = st[i] - ((i * 16 + grp) < BUCKET ? base_prev : base_main);
          ssq[i] = __ldg(sq + pos);
          valid[i] = p.mask != nullptr ? __ldg(p.mask + static_cast<int64_t>(b) * p.T + pos) : uint8_t(1);
        }
      }
      if (lt == 0) RTTS_STAMP(1, n, 1);
      mbar_wait(kv_free + st_i, ((n / kStages) & 1) ^ 1);       // PV of the tile that used this stage has completed
      if (lt == 0) RTTS_STAMP(1, n, 2);
      const __nv_bfloat16* qk_b = p.qk + static_cast<int64_t>(b) * p.T * p.ld + h * kDh + c * 8;
      const __nv_bfloat16* v_b = p.v + static_cast<int64_t>(b) * p.T * p.ld + h * kDh + c * 8;
#pragma unroll
      for (int i = 0; i < kPasses; ++i) {
        const int j = i * 16 + grp;
        const int64_t off = static_cast<int64_t>(st[i] - (j < BUCKET ? base_prev : base_main)) * p.ld;
        const uint32_t so = sw128_offset(j, c);
        cp_async16(sKP + so, qk_b + off);
        cp_async16(sV + so, v_b + off);
      }
      cp_async_commit();
      if (c == 0) {
#pragma unroll
        for (int i = 0; i < kPasses; ++i) {
          const int j = i * 16 + grp;
          const int pos = st[i] - (j < BUCKET ? base_prev : base_main);
          float inv;
          if (p.key_norm == RTTS_KEYNORM_L2) inv = 1.f / fmaxf(sqrtf(ssq[i]), 1e-12f);
          else inv = rsqrtf(ssq[i] * (1.f / kDh) + 1e-6f) * 0.125f;   // 1/sqrt(64)
          key_scale[j] = inv * p.score_scale_log2;
          key_pos[j] = valid[i] ? pos : (pos | kPadFlag);
          if (j >= kQOff) q_slot[j - kQOff] = st[i];
        }
      }
      cp_async_wait<0>();
      fence_proxy_async_smem();     // cp.async / st.shared data -> visible to the tensor-core (async) proxy
      mbar_arrive(kv_full + st_i);
      if (lt == 0) RTTS_STAMP(1, n, 3);
    }
  } else {
    // ================================================= softmax groups =============================================
    const int wg = warp >> 2;                       // 0 | 1
    const int m = tid - wg * 128;                   // query row = TMEM lane
    const uint32_t t_row = tmem + wg * 256 + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const int win0 = (m / BUCKET) * BUCKET;          // first key-tile row of this query's window
    const float mv = p.mask_value_log2, sv = p.self_value_log2;
    int n = wg;
    for (int tile = blockIdx.x + wg * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, n += 2) {
      const uint32_t ph = (n >> 1) & 1;
      const int st_i = n % kStages;
      uint8_t* stage = smem + st_i * L::kStageBytes;
      const float* key_scale = reinterpret_cast<const float*>(stage + L::kOffScale);
      const int* key_pos = reinterpret_cast<const int*>(stage + L::kOffPos);
      uint8_t* p_row_base = stage + L::kOffKP;
      const int row_bh = tile / p.tiles_per_row;
      if (m == 0) RTTS_STAMP(2, n, 0);
      mbar_wait(kv_full + st_i, (n / kStages) & 1);  // metadata of this stage is visible
      const int q_enc = key_pos[kQOff + m];
      int q_limit = p.causal ? (q_enc & ~kPadFlag) : (kPadFlag - 1);
      if (p.mask_mode == RTTS_MASK_QUERY_AND_KEY && (q_enc & kPadFlag)) q_limit = -1;   // padded query: all masked
      const int my_slot = reinterpret_cast<const int*>(stage + L::kOffSlot)[m];
      // stabiliser: |q_i| * score_scale * log2(e) = score_scale_log2^2 / key_scale[own row] for both key-norm variants ...
      // (key_scale = score_scale_log2 / |x| up to the norm's epsilon), times (1 + 2^-10) so rounding cannot push a score above it
      const float row_bound = p.score_scale_log2 * p.score_scale_log2 / key_scale[kQOff + m] * 1.001f;
      mbar_wait(s_full + wg, ph);
      tc_fence_after_sync();
      if (m == 0) RTTS_STAMP(2, n, 1);

      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      {
        // The query's own column sits in exactly one 32-column chunk per warp (warp-uniform); the same token can appear a second
        // time only in the look-back chunk of the first tile of a hash round.  Only those chunks pay for the self comparison, and
        // the position mask is skipped altogether when nothing can be masked (non-causal, no padding mask).
        const bool need_mask = p.causal || p.mask != nullptr;
        const int t_in_s = tile - row_bh * p.tiles_per_row;
        const bool round_start = (t_in_s * kQRows) % p.T == 0;
        const int diag_c0 = (kQOff + (m & ~31)) - win0;           // chunk holding columns of rows 32*(m/32) .. +31
        const uint32_t a_pos0 = smem_u32(key_pos + win0), a_scale0 = smem_u32(key_scale + win0);
        const uint32_t a_prow = smem_u32(p_row_base) + m * 128;
        const int m7 = m & 7;
        const float neg_bound = -row_bound;
#pragma unroll 1
        for (int c0 = 0; c0 < kWin; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(t_row + win0 + c0, r);
          const int col = win0 + c0;
          const uint32_t a_p = a_prow + (col >> 6) * (kQRows * 128);
          const int c16 = (col & 63) >> 3;
          const bool self_chunk = c0 == diag_c0 || (round_start && col < BUCKET);
          tmem_ld_wait();
          if (self_chunk) soft_chunk<true, true>(r, a_pos0 + c0 * 4, a_scale0 + c0 * 4, a_p, c16, m7, neg_bound, q_limit, q_enc, sum4);
          else if (need_mask) soft_chunk<true, false>(r, a_pos0 + c0 * 4, a_scale0 + c0 * 4, a_p, c16, m7, neg_bound, q_limit, q_enc, sum4);
          else soft_chunk<false, false>(r, a_pos0 + c0 * 4, a_scale0 + c0 * 4, a_p, c16, m7, neg_bound, q_limit, q_enc, sum4);
        }
      }
      float row_sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      if (m == 0) RTTS_STAMP(2, n, 5);
      float row_max = row_bound;
      bool redo = !(row_sum > 1e-30f);
      if (redo && row_bound < 60.f) {
        // Every term was exactly zero.  With bound < 60 a visible key cannot underflow (s - bound >= -2*bound > -126), so no key
        // is visible: only the query itself (masked to self_value) remains and the softmax is uniform over the self columns
        // (rp R8 "except when no other targets are available").  The query's own column is known; the same token can appear a
        // second time only when the look-back chunk comes from the previous hash round (first tile of a round).
        auto set_one = [&](int kc) {
          *reinterpret_cast<uint16_t*>(p_row_base + (kc >> 6) * (kQRows * 128) + sw128_offset(m, (kc & 63) >> 3) + (kc & 7) * 2) = 0x3F80;
        };
        set_one(kQOff + m);
        int n_self = 1;
        const int t_in = tile - row_bh * p.tiles_per_row;
        if ((t_in * kQRows) % p.T == 0 && win0 == 0) {       // window starts with the previous round's last chunk
          // branch-free count first: a duplicate is rare, the scan is not
          const uint32_t a_pos = smem_u32(key_pos);
          int n_dup = 0;
#pragma unroll 4
          for (int c4 = 0; c4 < BUCKET; c4 += 4) {
            const uint4 k4 = lds128(a_pos + c4 * 4);
            n_dup += (static_cast<int>(k4.x) == q_enc) + (static_cast<int>(k4.y) == q_enc) + (static_cast<int>(k4.z) == q_enc) +
                     (static_cast<int>(k4.w) == q_enc);
          }
          if (n_dup > 0) {
            for (int c = 0; c < BUCKET; ++c)
              if (key_pos[c] == q_enc) set_one(c);
            n_self += n_dup;
          }
        }
        row_sum = static_cast<float>(n_self);
        row_max = sv;
        redo = false;
      }
      if (__any_sync(0xffffffffu, redo)) {
        // exact two-pass arithmetic for the rows that need it; the TMEM loads are warp-collective, so every lane walks the loop
        float mx = -FLT_MAX;
#pragma unroll 1
        for (int c0 = 0; c0 < kWin; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(t_row + win0 + c0, r);
          tmem_ld_wait();
          if (redo) {
#pragma unroll 4
            for (int i = 0; i < 32; ++i) {
              const int kpi = key_pos[win0 + c0 + i];
              float sc = __uint_as_float(r[i]) * key_scale[win0 + c0 + i];
              sc = kpi > q_limit ? mv : sc;
              sc = kpi == q_enc ? sv : sc;
              mx = fmaxf(mx, sc);
            }
          }
        }
        float sm = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < kWin; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(t_row + win0 + c0, r);
          tmem_ld_wait();
          if (redo) {
#pragma unroll 1
            for (int q8 = 0; q8 < 4; ++q8) {
              float e8[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int col = win0 + c0 + q8 * 8 + i;
                const int kpi = key_pos[col];
                float sc = __uint_as_float(r[q8 * 8 + i]) * key_scale[col];
                sc = kpi > q_limit ? mv : sc;
                sc = kpi == q_enc ? sv : sc;
                e8[i] = exp2f(sc - mx);
                sm += e8[i];
              }
              const int col = win0 + c0 + q8 * 8;
              uint4 u;
              u.x = pack_bf16(e8[0], e8[1]); u.y = pack_bf16(e8[2], e8[3]);
              u.z = pack_bf16(e8[4], e8[5]); u.w = pack_bf16(e8[6], e8[7]);
              *reinterpret_cast<uint4*>(p_row_base + (col >> 6) * (kQRows * 128) + sw128_offset(m, (col & 63) >> 3)) = u;
            }
          }
        }
        if (redo) { row_max = mx; row_sum = sm; }
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
      tc_fence_before_sync();       // this thread's TMEM reads of S precede the MMAs that overwrite the region
      mbar_arrive(p_full + wg);
      if (m == 0) RTTS_STAMP(2, n, 2);

      mbar_wait(o_full + wg, ph);
      tc_fence_after_sync();
      if (m == 0) RTTS_STAMP(2, n, 3);
      {
        const float inv_sum = 1.f / row_sum;
        const int64_t slot = static_cast<int64_t>(row_bh) * RT + my_slot;
        uint4* dst = reinterpret_cast<uint4*>(p.o_rounds + slot * kDh);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r[32];
          tmem_ld32(t_row + kColO + half * 32, r);
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
      mbar_arrive(o_free + wg);     // O columns of this group may be overwritten
      if (m == 0) RTTS_STAMP(2, n, 4);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
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

static long long* g_fwd_trace = nullptr;   // debug only (rtts_debug_set_fwd_trace)

template <int BUCKET>
int launch_attn_fwd(const AttnFwdParams& p, int ctas, cudaStream_t stream) {
  using L = AttnFwdSmem<BUCKET>;
  static bool configured = false;   // idempotent attribute set; benign if raced
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(lsh_attn_fwd_kernel<BUCKET>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
    if (e != cudaSuccess) return fail(kErrCuda, "rtts_lsh_attn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int grid = ctas < kNumSMs ? ctas : kNumSMs;     // persistent: one CTA per SM
  lsh_attn_fwd_kernel<BUCKET><<<grid, kFwdThreads, L::kDynamic, stream>>>(p, ctas);
  return check_launch("rtts_lsh_attn_fwd");
}

}  // namespace rtts

using namespace rtts;

extern "C" int rtts_lsh_attn_fwd(const void* qk, const void* v, int64_t ld, const int32_t* sticker, const float* sumsq, const uint8_t* mask,
                                 const rtts_lsh_spec* spec, void* o_rounds, float* lse_rounds, int B, int T, int H, int dh,
                                 int R, int bucket, void* stream) {
  RTTS_REQUIRE(qk && v && sticker && sumsq && spec && o_rounds && lse_rounds, "rtts_lsh_attn_fwd: null pointer");
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
  p.sumsq = sumsq;
  p.mask = mask;
  p.o_rounds = static_cast<__nv_bfloat16*>(o_rounds);
  p.lse_rounds = lse_rounds;
  p.trace = g_fwd_trace;
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

// Debug hook (not part of the product ABI): device buffer of 3*32*8 int64 receiving clock64 stamps of CTA 0.
extern "C" void rtts_debug_set_fwd_trace(void* device_buffer) { g_fwd_trace = static_cast<long long*>(device_buffer); }
